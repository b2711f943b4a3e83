// STAND-IN for include/scene/camera.h:16-95 (fields the render backend reads).
#pragma once
#include <glm/glm.hpp>
namespace oka
{
class Camera
{
public:
    float fov = 45.0f;
    float znear = 0.1f, zfar = 1000.0f;
    glm::quat mOrientation = { 1.0f, 0.0f, 0.0f, 0.0f };
    glm::float3 position = { 0.0f, 0.0f, 10.0f };
    struct Matrices { glm::float4x4 perspective, invPerspective, view; } matrices;
    // camera.cpp:10-23 (first person): view = R(q) * T(-position)
    void updateViewMatrix()
    {
        const float w = mOrientation.w, x = mOrientation.x, y = mOrientation.y, z = mOrientation.z;
        const float r[3][3] = { { 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w) },
                                { 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w) },
                                { 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y) } };
        glm::mat4 v(1.0f);
        for (int row = 0; row < 3; ++row)
        {
            for (int col = 0; col < 3; ++col)
                v[col][row] = r[row][col];
            v[3][row] = -(r[row][0] * position.x + r[row][1] * position.y + r[row][2] * position.z);
        }
        matrices.view = v;
    }
    void updateAspectRatio(float) {} // the backend derives the projection itself (OptixRender.cpp:895-897)
};
} // namespace oka
