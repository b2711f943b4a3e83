"""Sample-sharded rendering behind the C ABI (sb_comm_* / sb_render_sharded): the library's own NCCL
communicator sums the accumulation buffers S.  The two-rank test needs two GPUs (one thread per rank in this process;
skipped on a one-GPU box -- run it with `gpurun --gpus 2`)."""
import threading

import numpy as np
import pytest

from strelka_b200 import BufferDesc, BufferFormat, RenderFactory, RenderType, SharedContext
from strelka_b200.scenes import make_cornell

pytestmark = pytest.mark.gpu


def _single(scene, settings, w, h, n):
    r = RenderFactory.createRender(RenderType.eCompute)
    r.setScene(scene)
    r.setSharedContext(SharedContext(mSettingsManager=settings))
    r.init()
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    r.render_iterations(buf, n)
    img = buf.map().copy()
    buf.destroy()
    r.destroy()
    return img


def test_world_of_one_is_the_plain_render():
    w = h = 96
    scene, st, _ = make_cornell(w, h, 12)
    want = _single(scene, st, w, h, 12)
    scene2, st2, _ = make_cornell(w, h, 12)
    r = RenderFactory.createRender(RenderType.eCompute)
    r.setScene(scene2)
    r.setSharedContext(SharedContext(mSettingsManager=st2))
    r.init()
    r.comm_init(r.comm_unique_id(), 0, 1)
    assert r.comm_world() == 1
    buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
    r.render_sharded(buf, 5)
    r.render_sharded(buf, 7)  # repeated calls keep accumulating (S stays this rank's partial sum)
    got = buf.map().copy()
    assert np.array_equal(got, want)
    r.comm_destroy()
    buf.destroy()
    r.destroy()


@pytest.mark.parametrize("nvls", ["1", "0"], ids=["fused-nvls-kernel", "nccl-allreduce"])
def test_two_ranks_sum_to_the_single_gpu_image(nvls, monkeypatch):
    """both exchange paths of sb_render_sharded: the fused NVLS all-reduce + resolve kernel (where NCCL >= 2.28 and the
    hardware offer it) and ncclAllReduce + k_resolve (forced by STRELKA_B200_NVLS=0)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("STRELKA_B200_NVLS", nvls)
    w = h = 128
    total = 13  # odd on purpose: the ranks own 7 and 6 samples
    scene, st, _ = make_cornell(w, h, total)
    want = _single(scene, st, w, h, total)
    uid = RenderFactory.createRender(RenderType.eCompute)  # only to reach sb_comm_get_unique_id
    uid.init()
    unique = uid.comm_unique_id()
    uid.destroy()
    out, errs = [None, None], []

    def rank(i):
        try:
            sc, s2, _ = make_cornell(w, h, total)
            r = RenderFactory.createRender(RenderType.eCompute, device=i)
            r.setScene(sc)
            r.setSharedContext(SharedContext(mSettingsManager=s2))
            r.init()
            r.comm_init(unique, i, 2)
            buf = r.createBuffer(BufferDesc(w, h, BufferFormat.FLOAT4))
            r.render_sharded(buf, 4)   # progressive: 4 + 4 + more than is left
            r.render_sharded(buf, 4)
            mid = buf.map().copy()
            r.render_sharded(buf, 100)
            c = r.counters()
            out[i] = (mid, buf.map().copy(), r.getSharedContext().mSubframeIndex, r.comm_exchange_path(), c["exchange_nvls"], c["exchange_ms"])
            r.comm_destroy()
            buf.destroy()
            r.destroy()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=rank, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert not errs, errs
    print("exchange path:", out[0][3], "| exchange_ms", out[0][5])
    assert out[0][3] == out[1][3] and out[0][4] == out[1][4]
    if nvls == "0":
        assert out[0][4] == 0 and "ncclAllReduce" in out[0][3]
    assert out[0][2] == 7 and out[1][2] == 6
    for i in range(2):
        np.testing.assert_allclose(out[i][1][..., :3], want[..., :3], rtol=2e-5, atol=1e-7)
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][0], out[1][0])
    # after 4 + 4 iterations per rank the ranks already hold all their samples (7 and 6 of 13): the intermediate image
    # is the final one
    np.testing.assert_allclose(out[0][0][..., :3], want[..., :3], rtol=2e-5, atol=1e-7)
