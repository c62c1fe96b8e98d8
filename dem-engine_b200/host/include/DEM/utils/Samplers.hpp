// DEM/utils/Samplers.hpp -- point samplers demo scripts use to generate input
// (counterpart of src/DEM/utils/Samplers.hpp of the reference: PDSampler :271-467, HCPSampler :498-533,
// GridSampler :536-573, DEMCylSurfSampler :616-641).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../HostSideHelpers.hpp"

namespace deme {

/// One process-wide random engine, and a draw from a normal distribution restricted to [minVal, maxVal] (what scripts use
/// for polydisperse sizes; Samplers.hpp:48-71 of the reference)
inline std::default_random_engine& rengine() {
    static std::default_random_engine engine;
    return engine;
}
template <typename T>
inline T sampleTruncatedDist(std::normal_distribution<T>& distribution, T minVal, T maxVal) {
    for (;;) {
        const T val = distribution(rengine());
        if (!(val < minVal) && !(val > maxVal)) return val;
    }
}

enum class SamplingType { REGULAR_GRID, POISSON_DISK, HCP_PACK };

class Sampler {
  public:
    explicit Sampler(float separation) : m_separation(separation) {}
    virtual ~Sampler() {}
    /// Points in the ball of given radius
    std::vector<float3> SampleSphere(const float3& center, float radius) {
        m_center = center;
        m_size = make_float3(radius, radius, radius);
        return Sample(4);
    }
    // {x, y, z} forms returning {{x, y, z}, ...} (what the reference's Python layer calls)
    std::vector<std::vector<float>> SampleBox(const std::vector<float>& center, const std::vector<float>& halfDim) {
        return Real3VectorToVecOfVec(SampleBox(xyz(center, "SampleBox"), xyz(halfDim, "SampleBox")));
    }
    std::vector<std::vector<float>> SampleSphere(const std::vector<float>& center, float radius) {
        return Real3VectorToVecOfVec(SampleSphere(xyz(center, "SampleSphere"), radius));
    }
    std::vector<std::vector<float>> SampleCylinderX(const std::vector<float>& center, float radius, float halfHeight) {
        return Real3VectorToVecOfVec(SampleCylinderX(xyz(center, "SampleCylinderX"), radius, halfHeight));
    }
    std::vector<std::vector<float>> SampleCylinderY(const std::vector<float>& center, float radius, float halfHeight) {
        return Real3VectorToVecOfVec(SampleCylinderY(xyz(center, "SampleCylinderY"), radius, halfHeight));
    }
    std::vector<std::vector<float>> SampleCylinderZ(const std::vector<float>& center, float radius, float halfHeight) {
        return Real3VectorToVecOfVec(SampleCylinderZ(xyz(center, "SampleCylinderZ"), radius, halfHeight));
    }
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
        if (volume == 4) return dot(v, v) <= m_size.x * m_size.x;
        if (volume == 2) return (v.y * v.y + v.z * v.z <= m_size.y * m_size.y) && std::fabs(v.x) <= m_size.x + fuzz;
        if (volume == 3) return (v.x * v.x + v.z * v.z <= m_size.x * m_size.x) && std::fabs(v.y) <= m_size.y + fuzz;
        return (v.x * v.x + v.y * v.y <= m_size.x * m_size.x) && std::fabs(v.z) <= m_size.z + fuzz;
    }
    static float3 xyz(const std::vector<float>& v, const char* who) {
        if (v.size() != 3) throw std::runtime_error(std::string(who) + ": a 3-element vector is expected");
        return make_float3(v[0], v[1], v[2]);
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
// Poisson-disk sampler: a maximal random set of points no two of which are closer than the separation (what
// DEMdemo_BallDrop / Repose / Centrifuge use to fill a volume without a lattice).  Dart throwing with a background grid
// (cell = separation / sqrt(3), so a cell holds at most one point): every accepted point spawns up to
// `pointsPerIteration` candidates in the shell [separation, 2 separation] around it; a candidate is kept if it lies in
// the volume and no point of the 5x5x5 cell neighbourhood is too close.  Seeded with 0 like the reference's
// (Samplers.hpp:279-281), so a script generates the same cloud every run.
class PDSampler : public Sampler {
  public:
    explicit PDSampler(float separation, int pointsPerIteration = 30)
        : Sampler(separation), m_ppi(pointsPerIteration), m_rng(0) {}
    void SetRandomEngineSeed(unsigned int seed) { m_rng.seed(seed); }

  protected:
    std::vector<float3> Sample(int volume) override {
        std::vector<float3> pts;
        const float r = m_separation;
        if (!(r > 0.f)) return pts;
        const float cell = r / std::sqrt(3.0f);
        // bounding box of the volume (half extents per axis)
        const float3 half = (volume == 1)   ? make_float3(m_size.x, m_size.x, m_size.z)
                            : (volume == 2) ? make_float3(m_size.x, m_size.y, m_size.y)
                            : (volume == 3) ? make_float3(m_size.x, m_size.y, m_size.x)
                                            : m_size;
        const float3 lo = m_center - half;
        auto key = [&](const float3& p) -> uint64_t {
            const uint64_t ix = (uint64_t)std::floor((p.x - lo.x) / cell + 2.f), iy = (uint64_t)std::floor((p.y - lo.y) / cell + 2.f),
                           iz = (uint64_t)std::floor((p.z - lo.z) / cell + 2.f);
            return ix | (iy << 21) | (iz << 42);
        };
        std::unordered_map<uint64_t, uint32_t> grid;
        auto too_close = [&](const float3& p) {
            const uint64_t k = key(p);
            const int64_t ix = (int64_t)(k & 0x1fffff), iy = (int64_t)((k >> 21) & 0x1fffff), iz = (int64_t)(k >> 42);
            for (int64_t dz = -2; dz <= 2; dz++)
                for (int64_t dy = -2; dy <= 2; dy++)
                    for (int64_t dx = -2; dx <= 2; dx++) {
                        const int64_t x = ix + dx, y = iy + dy, z = iz + dz;
                        if (x < 0 || y < 0 || z < 0) continue;
                        auto it = grid.find((uint64_t)x | ((uint64_t)y << 21) | ((uint64_t)z << 42));
                        if (it == grid.end()) continue;
                        const float3 d = pts[it->second] - p;
                        if (dot(d, d) < r * r) return true;
                    }
            return false;
        };
        std::uniform_real_distribution<float> U(0.f, 1.f);
        // first point: the centre of the volume (always inside)
        pts.push_back(m_center);
        grid[key(m_center)] = 0;
        std::vector<uint32_t> active(1, 0u);
        while (!active.empty()) {
            const size_t pick = (size_t)(U(m_rng) * (float)active.size()) % active.size();
            const float3 base = pts[active[pick]];
            bool spawned = false;
            for (int t = 0; t < m_ppi; t++) {
                // uniform direction, radius uniform in volume over the shell [r, 2r]
                const float cz = 2.f * U(m_rng) - 1.f, phi = 6.2831853f * U(m_rng);
                const float sz = std::sqrt(std::max(0.f, 1.f - cz * cz));
                const float rad = r * std::cbrt(1.f + 7.f * U(m_rng));
                const float3 c = base + make_float3(sz * std::cos(phi), sz * std::sin(phi), cz) * rad;
                if (!accept(volume, c) || too_close(c)) continue;
                grid[key(c)] = (uint32_t)pts.size();
                active.push_back((uint32_t)pts.size());
                pts.push_back(c);
                spawned = true;
            }
            if (!spawned) {
                active[pick] = active.back();
                active.pop_back();
            }
        }
        return pts;
    }

  private:
    int m_ppi;
    std::mt19937 m_rng;
};

/// Points on the mantle of a cylinder, rows along the axis spaced `spacing * ParticleRad` apart (a shell of particles
/// that resembles a cylindrical surface; Samplers.hpp:616-641 of the reference)
inline std::vector<float3> DEMCylSurfSampler(float3 CylCenter, float3 CylAxis, float CylRad, float CylHeight,
                                             float ParticleRad, float spacing = 1.2f) {
    std::vector<float3> points;
    const float step = spacing * ParticleRad;
    const unsigned int rows = (unsigned int)(2.0 * 3.14159265358979323846 * CylRad / step);
    if (rows == 0 || !(step > 0.f)) return points;
    const float3 a = normalize(CylAxis);
    // any unit vector perpendicular to the axis, and the one completing the frame
    const float3 helper = (std::fabs(a.x) < 0.9f) ? make_float3(1, 0, 0) : make_float3(0, 1, 0);
    const float3 u = normalize(cross(a, helper)), v = cross(a, u);
    for (unsigned int i = 0; i < rows; i++) {
        const float ang = 6.28318530717958647692f * (float)i / (float)rows;
        const float3 radial = u * std::cos(ang) + v * std::sin(ang);
        const float3 start = CylCenter + a * (CylHeight / 2.f) + radial * CylRad;
        for (float d = 0.f; d <= CylHeight; d += step) points.push_back(start - a * d);
    }
    return points;
}

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
// {x, y, z} forms of the three free samplers (Samplers.hpp:589-660 of the reference)
namespace sampler_detail {
inline float3 xyz(const std::vector<float>& v, const char* who) {
    if (v.size() != 3) throw std::runtime_error(std::string(who) + ": a 3-element vector is expected");
    return make_float3(v[0], v[1], v[2]);
}
}  // namespace sampler_detail
inline std::vector<float3> DEMBoxGridSampler(const std::vector<float>& BoxCenter, const std::vector<float>& HalfDims,
                                             float GridSizeX, float GridSizeY = -1.0, float GridSizeZ = -1.0) {
    return DEMBoxGridSampler(sampler_detail::xyz(BoxCenter, "DEMBoxGridSampler"), sampler_detail::xyz(HalfDims, "DEMBoxGridSampler"),
                             GridSizeX, GridSizeY, GridSizeZ);
}
inline std::vector<float3> DEMBoxHCPSampler(const std::vector<float>& BoxCenter, const std::vector<float>& HalfDims, float GridSize) {
    return DEMBoxHCPSampler(sampler_detail::xyz(BoxCenter, "DEMBoxHCPSampler"), sampler_detail::xyz(HalfDims, "DEMBoxHCPSampler"), GridSize);
}
inline std::vector<std::vector<float>> DEMCylSurfSampler(const std::vector<float>& CylCenter, const std::vector<float>& CylAxis,
                                                         float CylRad, float CylHeight, float ParticleRad, float spacing = 1.2f) {
    return Real3VectorToVecOfVec(DEMCylSurfSampler(sampler_detail::xyz(CylCenter, "DEMCylSurfSampler"),
                                                   sampler_detail::xyz(CylAxis, "DEMCylSurfSampler"), CylRad, CylHeight, ParticleRad, spacing));
}

/// A box filled layer by layer: Poisson-disk sampling in horizontal planes `padding_factor * diam` apart -- much cheaper
/// than a volumetric Poisson-disk cloud, at the price of regular spacing across the layers (Samplers.hpp:464-496)
inline std::vector<float3> PDLayerSampler_BOX(float3 center, float3 hdims, float diam, float padding_factor = 1.02f,
                                              bool verbose = false) {
    const float pitch = diam * padding_factor, top = center.z + hdims.z;
    std::vector<float3> all;
    if (!(pitch > 0.f)) return all;
    PDSampler sampler(pitch);
    const float3 flat = make_float3(hdims.x, hdims.y, 0.f);
    for (float z = center.z - hdims.z; z < top; z += pitch) {
        if (verbose) std::printf("Create layer at %g\n", z);
        const std::vector<float3> layer = sampler.SampleBox(make_float3(center.x, center.y, z), flat);
        all.insert(all.end(), layer.begin(), layer.end());
    }
    return all;
}

}  // namespace deme
