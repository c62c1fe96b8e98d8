// Host-only check of the facade's checkpoint readers (no GPU needed: only static members are used).
// usage: csv_readers_check <clumps.csv> <contacts.csv>
#include <DEM/API.h>

#include <cstdio>

using namespace deme;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string clumps = argv[1], contacts = argv[2];
    auto xyz = DEMSolver::ReadClumpXyzFromCsv(clumps);
    auto quat = DEMSolver::ReadClumpQuatFromCsv(clumps);
    auto vel = DEMSolver::ReadClumpVelFromCsv(clumps);
    auto angvel = DEMSolver::ReadClumpAngVelFromCsv(clumps);
    for (const auto& kv : xyz) {
        printf("type %s n %zu\n", kv.first.c_str(), kv.second.size());
        for (size_t i = 0; i < kv.second.size(); i++) {
            const float3 p = kv.second[i], v = vel[kv.first][i], w = angvel[kv.first][i];
            const float4 q = quat[kv.first][i];
            printf("clump %s %zu %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", kv.first.c_str(), i,
                   p.x, p.y, p.z, q.w, q.x, q.y, q.z, v.x, v.y, v.z, w.x, w.y, w.z);
        }
    }
    auto pairs = DEMSolver::ReadContactPairsFromCsv(contacts);
    auto wc = DEMSolver::ReadContactWildcardsFromCsv(contacts);
    printf("pairs %zu wildcards %zu\n", pairs.size(), wc.size());
    for (size_t i = 0; i < pairs.size(); i++)
        printf("pair %u %u %.9g %.9g %.9g %.9g\n", pairs[i].first, pairs[i].second, wc["delta_tan_x"][i], wc["delta_tan_y"][i],
               wc["delta_tan_z"][i], wc["delta_time"][i]);
    auto sa = DEMSolver::ReadContactPairsFromCsv(contacts, "SA");
    printf("sa_pairs %zu\n", sa.size());
    return 0;
}
