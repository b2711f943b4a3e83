"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank renders its sample stride with the host
instantiation of the device code (tests/emul), S is all-reduced through strelka_b200.distributed and
resolved; rank 0 compares with the single-rank render of all samples."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import pyemul
    from strelka_b200.distributed import allreduce_accumulation, local_sample_count, shard_settings
    from strelka_b200.scenes import make_cornell

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = h = 32
    spp = 9  # odd on purpose: ranks get 5 and 4 samples
    scene, settings, _ = make_cornell(w, h, spp)
    e = pyemul.EmulScene(scene)
    shard_settings(settings, rank, world)
    n_local = local_sample_count(spp, rank, world)
    _, S, cnt = e.render(settings, w, h, n_local, chunk_max=3)
    t = torch.from_numpy(S.reshape(-1))
    allreduce_accumulation(t)
    counts = torch.tensor([float(n_local), float(cnt["radiance_rays"])], dtype=torch.float64)
    dist.all_reduce(counts)
    if rank == 0:
        np.save(out_path, np.concatenate([t.numpy(), counts.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import pyemul
    from oracle import pyoracle
    from strelka_b200.scenes import make_cornell

    out = str(tmp_path / "S.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    S2, n_total, rays_total = got[:-2], got[-2], got[-1]
    w = h = 32
    scene, settings, _ = make_cornell(w, h, 9)
    _, S1, cnt = pyemul.EmulScene(scene).render(settings, w, h, 9, chunk_max=9)
    assert n_total == 9 and rays_total == cnt["radiance_rays"]
    np.testing.assert_allclose(S2, S1.reshape(-1), rtol=3e-6, atol=1e-9)
    # and the resolved image equals the oracle's single-process render within the parity gate
    e = pyoracle.exposure(settings)
    A = S2.reshape(h, w, 4)[..., :3] / 9.0
    img = A / (e - A * e)
    ref, _, _, _ = pyoracle.OracleScene(scene).render(settings, w, h, 9)
    err = np.sqrt(np.mean((img - ref[..., :3]) ** 2) / np.mean(ref[..., :3] ** 2))
    assert err < 1e-5


def test_local_sample_count_partitions_all_samples():
    from strelka_b200.distributed import local_sample_count

    for total in (1, 7, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            assert sum(local_sample_count(total, r, world) for r in range(world)) == total
