// STAND-IN for include/scene/scene.h (the arrays + getters a render backend consumes, scene.h:21-155,
// 229-324).  Same member names and layouts as the reference.
#pragma once
#include "camera.h"
#include <cstdint>
#include <string>
#include <vector>
#include <materialmanager/materialmanager.h>
namespace oka
{
struct Mesh { uint32_t mIndex, mCount, mVbOffset, mVertexCount; };
struct Curve
{
    enum class Type : uint8_t { eLinear, eCubic };
    uint32_t mVertexCountsStart, mVertexCountsCount, mPointsStart, mPointsCount, mWidthsStart, mWidthsCount;
};
struct Instance
{
    glm::mat4 transform;
    enum class Type : uint8_t { eMesh, eLight, eCurve } type;
    union { uint32_t mMeshId; uint32_t mCurveId; };
    uint32_t mMaterialId = 0;
    uint32_t mLightId = (uint32_t)-1;
};
class Scene
{
public:
    struct MaterialDescription
    {
        enum class Type { eMdl, eMaterialX } type;
        std::string code, file, name;
        bool hasColor = false;
        glm::float3 color;
        std::vector<MaterialManager::Param> params;
    };
    struct Vertex { glm::float3 pos; uint32_t tangent; uint32_t normal; uint32_t uv; float pad0; float pad1; };
    struct Light { glm::float4 points[4]; glm::float4 color = glm::float4(1, 1, 1, 1); glm::float4 normal; int type; float halfAngle; float pad0; float pad1; };
    std::vector<Vertex>& getVertices() { return mVertices; }
    std::vector<uint32_t>& getIndices() { return mIndices; }
    std::vector<MaterialDescription>& getMaterials() { return mMaterialsDescs; }
    std::vector<Light>& getLights() { return mLights; }
    Camera& getCamera(uint32_t index) { if (mCameras.empty()) mCameras.emplace_back(); return mCameras[index]; }
    const std::vector<Instance>& getInstances() const { return mInstances; }
    const std::vector<Mesh>& getMeshes() const { return mMeshes; }
    const std::vector<Curve>& getCurves() const { return mCurves; }
    const std::vector<glm::float3>& getCurvesPoint() const { return mCurvePoints; }
    const std::vector<float>& getCurvesWidths() const { return mCurveWidths; }
    const std::vector<uint32_t>& getCurvesVertexCounts() const { return mCurveVertexCounts; }
    uint32_t addMaterial(const MaterialDescription& m) { mMaterialsDescs.push_back(m); return uint32_t(mMaterialsDescs.size() - 1); }
    // public in the shim so that tests can fill them directly
    std::vector<Vertex> mVertices;
    std::vector<uint32_t> mIndices;
    std::vector<Mesh> mMeshes;
    std::vector<Curve> mCurves;
    std::vector<glm::float3> mCurvePoints;
    std::vector<float> mCurveWidths;
    std::vector<uint32_t> mCurveVertexCounts;
    std::vector<Instance> mInstances;
    std::vector<Light> mLights;
    std::vector<MaterialDescription> mMaterialsDescs;
    std::vector<Camera> mCameras;
};
} // namespace oka
