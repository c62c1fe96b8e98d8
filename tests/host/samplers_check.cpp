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
    return 0;
}
