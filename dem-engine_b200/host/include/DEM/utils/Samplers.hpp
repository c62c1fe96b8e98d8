// DEM/utils/Samplers.hpp -- point samplers demo scripts use to generate input
// (counterpart of src/DEM/utils/Samplers.hpp of the reference: HCPSampler :498-533, GridSampler :536-573).
#pragma once
#include <cmath>
#include <vector>

#include "../HostSideHelpers.hpp"

namespace deme {

class Sampler {
  public:
    explicit Sampler(float separation) : m_separation(separation) {}
    virtual ~Sampler() {}
    /// Points in the box centred at `center` with half dimensions `halfDim`
    std::vector<float3> SampleBox(const float3& center, const float3& halfDim) {
        m_center = center;
        m_size = halfDim;
        return Sample(0);
    }
    /// Points in the z-aligned cylinder of given radius and half height
    std::vector<float3> SampleCylinderZ(const float3& center, float radius, float halfHeight) {
        m_center = center;
        m_size = make_float3(radius, radius, halfHeight);
        return Sample(1);
    }
    /// Points in the x- / y-aligned cylinder of given radius and half height
    std::vector<float3> SampleCylinderX(const float3& center, float radius, float halfHeight) {
        m_center = center;
        m_size = make_float3(halfHeight, radius, radius);
        return Sample(2);
    }
    std::vector<float3> SampleCylinderY(const float3& center, float radius, float halfHeight) {
        m_center = center;
        m_size = make_float3(radius, halfHeight, radius);
        return Sample(3);
    }
    virtual void SetSeparation(float separation) { m_separation = separation; }
    virtual float GetSeparation() const { return m_separation; }

  protected:
    virtual std::vector<float3> Sample(int volume) = 0;
    bool accept(int volume, const float3& p) const {
        const float3 v = p - m_center;
        const float fuzz = (m_size.x < 1) ? 1e-6f * m_size.x : 1e-6f;
        if (volume == 0)
            return std::fabs(v.x) <= m_size.x + fuzz && std::fabs(v.y) <= m_size.y + fuzz && std::fabs(v.z) <= m_size.z + fuzz;
        if (volume == 2) return (v.y * v.y + v.z * v.z <= m_size.y * m_size.y) && std::fabs(v.x) <= m_size.x + fuzz;
        if (volume == 3) return (v.x * v.x + v.z * v.z <= m_size.x * m_size.x) && std::fabs(v.y) <= m_size.y + fuzz;
        return (v.x * v.x + v.y * v.y <= m_size.x * m_size.x) && std::fabs(v.z) <= m_size.z + fuzz;
    }
    float m_separation;
    float3 m_center = make_float3(0, 0, 0);
    float3 m_size = make_float3(0, 0, 0);
};

class HCPSampler : public Sampler {
  public:
    explicit HCPSampler(float separation) : Sampler(separation) {}

  private:
    std::vector<float3> Sample(int t) override {
        std::vector<float3> out;
        const float3 bl = m_center - m_size;
        const float dx = m_separation;
        const float dy = m_separation * (float)(std::sqrt(3.0) / 2);
        const float dz = m_separation * (float)(std::sqrt(2.0 / 3.0));
        const int nx = (int)(2 * m_size.x / dx) + 1, ny = (int)(2 * m_size.y / dy) + 1, nz = (int)(2 * m_size.z / dz) + 1;
        for (int k = 0; k < nz; k++) {
            const float offset_y = (k % 2 == 0) ? 0 : dy / 3;
            for (int j = 0; j < ny; j++) {
                const float offset_x = ((j + k) % 2 == 0) ? 0 : dx / 2;
                for (int i = 0; i < nx; i++) {
                    const float3 p = bl + make_float3(offset_x + i * dx, offset_y + j * dy, k * dz);
                    if (accept(t, p)) out.push_back(p);
                }
            }
        }
        return out;
    }
};

class GridSampler : public Sampler {
  public:
    explicit GridSampler(float separation) : Sampler(separation), m_sep3D(make_float3(separation, separation, separation)) {}
    explicit GridSampler(const float3& separation) : Sampler(separation.x), m_sep3D(separation) {}
    void SetSeparation(float separation) override { m_sep3D = make_float3(separation, separation, separation); }

  private:
    std::vector<float3> Sample(int t) override {
        std::vector<float3> out;
        const float3 bl = m_center - m_size;
        const int nx = (int)(2 * m_size.x / m_sep3D.x) + 1, ny = (int)(2 * m_size.y / m_sep3D.y) + 1,
                  nz = (int)(2 * m_size.z / m_sep3D.z) + 1;
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < ny; j++)
                for (int k = 0; k < nz; k++) {
                    const float3 p = bl + make_float3(i * m_sep3D.x, j * m_sep3D.y, k * m_sep3D.z);
                    if (accept(t, p)) out.push_back(p);
                }
        return out;
    }
    float3 m_sep3D;
};

/// DEMBoxGridSampler(BoxCenter, HalfDims, GridSizeX, GridSizeY, GridSizeZ)
inline std::vector<float3> DEMBoxGridSampler(float3 BoxCenter, float3 HalfDims, float GridSizeX, float GridSizeY = -1.0,
                                             float GridSizeZ = -1.0) {
    if (GridSizeY < 0) GridSizeY = GridSizeX;
    if (GridSizeZ < 0) GridSizeZ = GridSizeX;
    GridSampler sampler(make_float3(GridSizeX, GridSizeY, GridSizeZ));
    return sampler.SampleBox(BoxCenter, HalfDims);
}
inline std::vector<float3> DEMBoxHCPSampler(float3 BoxCenter, float3 HalfDims, float GridSize) {
    HCPSampler sampler(GridSize);
    return sampler.SampleBox(BoxCenter, HalfDims);
}

}  // namespace deme
