// Host-only check of the point samplers (no GPU): prints the points of each sampler for the Python test to examine.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <cstdio>

using namespace deme;

static void dump(const char* name, const std::vector<float3>& p) {
    printf("%s %zu\n", name, p.size());
    for (const auto& q : p) printf("%.9g %.9g %.9g\n", q.x, q.y, q.z);
}

int main() {
    PDSampler pd(0.05f);
    dump("pd_box", pd.SampleBox(make_float3(0.1f, -0.2f, 0.3f), make_float3(0.5f, 0.4f, 0.3f)));
    PDSampler pd2(0.05f);
    dump("pd_box_again", pd2.SampleBox(make_float3(0.1f, -0.2f, 0.3f), make_float3(0.5f, 0.4f, 0.3f)));
    PDSampler pd3(0.04f);
    dump("pd_cylz", pd3.SampleCylinderZ(make_float3(0, 0, 0), 0.3f, 0.2f));
    HCPSampler hcp(0.05f);
    dump("hcp_box", hcp.SampleBox(make_float3(0, 0, 0), make_float3(0.3f, 0.3f, 0.3f)));
    dump("grid_box", DEMBoxGridSampler(make_float3(0, 0, 0), make_float3(0.3f, 0.3f, 0.3f), 0.05f));
    dump("cyl_surf", DEMCylSurfSampler(make_float3(0, 0, 1), make_float3(0, 0, 2), 0.5f, 1.0f, 0.02f, 1.2f));
    PDSampler pd4(0.04f);
    dump("pd_sphere", pd4.SampleSphere(make_float3(1, 2, 3), 0.3f));
    HCPSampler hcp2(0.05f);
    dump("hcp_sphere", hcp2.SampleSphere(make_float3(0, 0, 0), 0.3f));
    dump("pd_layers", PDLayerSampler_BOX(make_float3(0, 0, 0.5f), make_float3(0.3f, 0.2f, 0.1f), 0.04f, 1.05f));
    // the {x, y, z} forms give the same points
    std::vector<float3> via_vec;
    HCPSampler hcp3(0.05f);
    for (const auto& v : hcp3.SampleBox(std::vector<float>{0, 0, 0}, std::vector<float>{0.3f, 0.3f, 0.3f}))
        via_vec.push_back(make_float3(v[0], v[1], v[2]));
    dump("hcp_box_vec", via_vec);
    dump("grid_box_vec", DEMBoxGridSampler(std::vector<float>{0, 0, 0}, std::vector<float>{0.3f, 0.3f, 0.3f}, 0.05f));
    std::normal_distribution<float> dist(1.f, 0.5f);
    std::vector<float3> trunc;
    for (int i = 0; i < 300; i++) trunc.push_back(make_float3(sampleTruncatedDist<float>(dist, 0.8f, 1.2f), 0, 0));
    dump("truncated", trunc);
    return 0;
}
