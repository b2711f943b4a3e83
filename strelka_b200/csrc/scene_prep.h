// Host-side preparation of a scene for the device: validation of the borrowed arrays, the instance
// table (transforms, inverse transforms, visibility masks, material fallback) and the world-space
// primitive numbering.  Shared by the C ABI (sb_api.cu) and the CPU emulation harness of the tests.
// Reference: OptiXRender::createAccelerationStructure instance loop (OptixRender.cpp:410-441),
// createSbt per-instance records (:757-797), createCurve segment list (:232-245).
#pragma once
#include "bvh_build.h"
#include <stdexcept>
#include <string>
#include <vector>

namespace sb
{

struct ScenePrep
{
    std::vector<InstDev> inst;
    std::vector<uint32_t> triFirst; // numInstances + 1
    std::vector<SegInfo> segInfo; // world segments, instance-major
    uint64_t numTris = 0;
    uint32_t numMaterials = 1;
};

inline void prepare_scene(const sb_scene_view* v, ScenePrep& out)
{
    // ---- validate (the device code indexes these arrays unchecked) -------------------------------
    for (uint32_t m = 0; m < v->num_meshes; ++m)
    {
        const sb_mesh& mm = v->meshes[m];
        if (uint64_t(mm.index) + mm.count > v->num_indices || uint64_t(mm.vb_offset) + mm.vertex_count > v->num_vertices)
            throw std::runtime_error("sb_set_scene: mesh " + std::to_string(m) + " exceeds the index/vertex buffers");
        for (uint32_t k = 0; k < mm.count; ++k)
            if (v->indices[mm.index + k] >= mm.vertex_count)
                throw std::runtime_error("sb_set_scene: mesh " + std::to_string(m) + " has an index beyond its vertex count");
    }
    for (uint32_t cidx = 0; cidx < v->num_curves; ++cidx)
    {
        const sb_curve& cc = v->curves[cidx];
        if (uint64_t(cc.vertex_counts_start) + cc.vertex_counts_count > v->num_curve_vertex_counts ||
            uint64_t(cc.points_start) + cc.points_count > v->num_curve_points)
            throw std::runtime_error("sb_set_scene: curve " + std::to_string(cidx) + " exceeds the curve buffers");
        uint64_t total = 0;
        for (uint32_t k = 0; k < cc.vertex_counts_count; ++k)
            total += v->curve_vertex_counts[cc.vertex_counts_start + k];
        if (total > cc.points_count)
            throw std::runtime_error("sb_set_scene: curve " + std::to_string(cidx) + " vertex counts exceed its points");
    }
    // ---- instance table, world primitive offsets -----------------------------------------------------
    std::vector<InstDev>& inst = out.inst;
    inst.assign(v->num_instances, InstDev());
    std::vector<uint32_t>& triFirst = out.triFirst;
    triFirst.assign(size_t(v->num_instances) + 1, 0u);
    std::vector<SegInfo>& segInfo = out.segInfo;
    segInfo.clear();
    uint64_t& numTris = out.numTris;
    numTris = 0;
    const uint32_t numMaterials = std::max<uint32_t>(v->num_materials, 1u);
    out.numMaterials = numMaterials;
    for (uint32_t i = 0; i < v->num_instances; ++i)
    {
        const sb_instance& in = v->instances[i];
        InstDev& d = inst[i];
        std::memset(&d, 0, sizeof(d));
        for (int r = 0; r < 3; ++r)
            for (int col = 0; col < 4; ++col)
                d.o2w.m[r * 4 + col] = in.transform[col * 4 + r]; // glm column-major -> row-major (OptixRender.cpp:438)
        d.w2o = invert_affine(d.o2w);
        d.scale = affine_uniform_scale(d.o2w);
        d.type = in.type;
        d.geom = in.geom_id;
        d.material = (in.material_id == 0xffffffffu || in.material_id >= numMaterials) ? 0u : in.material_id; // OptixRender.cpp:766
        d.light = in.light_id;
        triFirst[i] = uint32_t(numTris);
        if (in.type == SB_INSTANCE_MESH || in.type == SB_INSTANCE_LIGHT)
        {
            d.mask = (in.type == SB_INSTANCE_MESH) ? kMaskTriangle : kMaskLight; // OptixRender.cpp:418-432
            if (in.geom_id >= v->num_meshes)
                throw std::runtime_error("sb_set_scene: instance " + std::to_string(i) + " references a missing mesh");
            d.numPrims = v->meshes[in.geom_id].count / 3;
            d.firstPrim = uint32_t(numTris);
            numTris += d.numPrims;
        }
        else if (in.type == SB_INSTANCE_CURVE)
        {
            d.mask = kMaskCurve;
            if (in.geom_id >= v->num_curves)
                throw std::runtime_error("sb_set_scene: instance " + std::to_string(i) + " references a missing curve");
            const sb_curve& cc = v->curves[in.geom_id];
            // segment list of OptiXRender::createCurve (OptixRender.cpp:232-245)
            uint32_t offsetInside = 0, prim = 0;
            d.firstPrim = uint32_t(segInfo.size());
            for (uint32_t ci = 0; ci < cc.vertex_counts_count; ++ci)
            {
                const uint32_t ncp = v->curve_vertex_counts[cc.vertex_counts_start + ci];
                const int nseg = int(ncp) - 3;
                for (int s = 0; s < nseg; ++s)
                {
                    SegInfo si;
                    si.prim = prim++;
                    si.inst = i;
                    si.firstPoint = cc.points_start + offsetInside + uint32_t(s);
                    si.span = 1u << 16;
                    segInfo.push_back(si);
                }
                offsetInside += ncp;
            }
            d.numPrims = prim;
        }
        else
        {
            throw std::runtime_error("sb_set_scene: instance " + std::to_string(i) + " has an unknown type");
        }
        if (numTris > 0x0fffffffull || segInfo.size() > 0x0fffffffull || v->num_instances > 0x0fffffffu)
            throw std::runtime_error("sb_set_scene: scene too large for 28-bit primitive/instance ids");
    }
    triFirst[v->num_instances] = uint32_t(numTris);

}

} // namespace sb
