// Particles are poured into a box batch by batch: AddClumps + UpdateClumps on a running simulation, the pattern of the
// reference's filling loops (DEMdemo_GRCPrep_Part1.cpp:140-153, DEMdemo_Hopper_Sphere_Cylinder.cpp:260-269:
// "AddClumps(...); UpdateClumps();").
// Prints, after every batch, the clump count, the height of the pile and how far the first batch has moved between
// the instant before and the instant after the update (it must not move at all: its state is carried over exactly).
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

using namespace deme;

int main(int argc, char** argv) {
    const int n_batches = argc > 1 ? std::atoi(argv[1]) : 3;
    DEMSolver DEMSim;
    DEMSim.SetVerbosity(QUIET);
    auto mat = DEMSim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.5}, {"mu", 0.4}, {"Crr", 0.0}});
    const float scale = 0.01f;
    // unit-size mass properties of the 3-sphere clump (DEMdemo_Mixer.cpp:68-72); Scale() takes them to the real size
    auto tmpl = DEMSim.LoadClumpType(2.6e3f * 5.5886717f, make_float3(2.928f, 2.6029f, 3.9908f) * 2.6e3f,
                                     GetDEMEDataFile("clumps/3_clump.csv").string(), mat);
    tmpl->Scale(scale);
    DEMSim.InstructBoxDomainDimension(0.4, 0.4, 1.2);
    DEMSim.InstructBoxDomainBoundingBC("top_open", mat);
    DEMSim.SetGravitationalAcceleration(make_float3(0, 0, -9.81));
    DEMSim.SetInitTimeStep(1e-5);
    DEMSim.SetCDUpdateFreq(10);
    DEMSim.UseAdaptiveUpdateFreq(false);  // (this script compares contact lists: keep the cycle length, hence the margin, fixed)

    HCPSampler sampler(3.f * scale);
    auto layer = [&](float z) { return sampler.SampleBox(make_float3(0, 0, z), make_float3(0.15f, 0.15f, 0.03f)); };
    auto first = DEMSim.AddClumps(tmpl, layer(-0.5f));
    first->SetVel(make_float3(0, 0, -0.5f));
    auto tracker = DEMSim.Track(first);
    auto max_z = DEMSim.CreateInspector("clump_max_z");
    auto kinetic = DEMSim.CreateInspector("clump_kinetic_energy");
    DEMSim.Initialize();
    const size_t n_first = first->GetNumClumps();
    for (int b = 1; b <= n_batches; b++) {
        DEMSim.DoDynamicsThenSync(0.05);
        const std::vector<float3> before = tracker->Positions();
        const std::vector<float3> vbefore = tracker->Velocities();
        const size_t contacts_before = DEMSim.GetNumContacts();
        auto more = DEMSim.AddClumps(tmpl, layer(max_z->GetValue() + 0.06f));
        more->SetVel(make_float3(0, 0, -0.5f));
        DEMSim.UpdateClumps();
        const std::vector<float3> after = tracker->Positions();
        const std::vector<float3> vafter = tracker->Velocities();
        double moved = 0., dv = 0.;
        for (size_t i = 0; i < n_first; i++) {
            moved = std::max<double>(moved, length(after[i] - before[i]));
            dv = std::max<double>(dv, length(vafter[i] - vbefore[i]));
        }
        printf("Batch %d: clumps = %zu, max z = %.5f, kinetic energy = %.6e, first batch moved = %.3e, dv = %.3e, "
               "contacts before/after = %zu/%zu, t = %.4f\n",
               b, DEMSim.GetNumClumps(), max_z->GetValue(), kinetic->GetValue(), moved, dv, contacts_before,
               DEMSim.GetNumContacts(), DEMSim.GetSimTime());
    }
    DEMSim.DoDynamicsThenSync(0.05);
    printf("Final: clumps = %zu, max z = %.5f, t = %.4f\n", DEMSim.GetNumClumps(), max_z->GetValue(), DEMSim.GetSimTime());
    // contact queries: every listed pair (owner ids, sorted by A), and the force pairs acting on the first batch
    const auto all_pairs = DEMSim.GetContacts();
    const auto clump_pairs = DEMSim.GetClumpContacts();
    bool sorted = true;
    for (size_t i = 1; i < all_pairs.size(); i++) sorted = sorted && all_pairs[i - 1].first <= all_pairs[i].first;
    std::vector<float3> points, forces;
    const size_t nf = tracker->GetContactForcesForAll(points, forces);
    float3 sum = make_float3(0, 0, 0);
    float zmin = 1e30f, zmax = -1e30f;
    for (size_t i = 0; i < nf; i++) {
        sum = sum + forces[i];
        zmin = std::min(zmin, points[i].z);
        zmax = std::max(zmax, points[i].z);
    }
    printf("Contacts: listed = %zu (GetNumContacts %zu), clump-clump = %zu, sorted = %d; force pairs on first batch = %zu, "
           "sum fz = %.4f, point z range = [%.4f, %.4f]\n",
           all_pairs.size(), DEMSim.GetNumContacts(), clump_pairs.size(), (int)sorted, nf, sum.z, zmin, zmax);
    // solver read-outs: broad-phase grid, margin, memory, neighbours of one clump, the simulated-time clock
    const double t_end = DEMSim.GetSimTime();
    DEMSim.SetSimTime(1.5);
    printf("Solver: bin size = %.6f, bins = %zu, margin = %.6f, device MB = %.1f, clump 0 touches %zu clumps, "
           "time %.4f -> %.4f\n", DEMSim.GetBinSize(), DEMSim.GetBinNum(), DEMSim.GetExpandFactor(),
           DEMSim.GetDeviceMemUsageDynamic() / 1048576.0, DEMSim.GetOwnerContactClumps(0).size(), t_end, DEMSim.GetSimTime());
    DEMSim.WriteContactFileIncludingPotentialPairs("DemoOutput_FillInBatches_contacts.csv");
    // freeze the bottom of the pile: clumps below a height join a fixed family, and stop moving
    DEMSim.SetFamilyFixed(5);
    const size_t frozen = DEMSim.ChangeClumpFamily(5, {-1., 1.}, {-1., 1.}, {-1., -0.55});
    DEMSim.DoDynamicsThenSync(0.002);
    float vmax_frozen = 0.f;
    size_t in_family = 0;
    const std::vector<float3> vel_all = tracker->Velocities();
    for (size_t i = 0; i < n_first; i++)
        if (tracker->GetFamily(i) == 5) {
            in_family++;
            vmax_frozen = std::max(vmax_frozen, length(vel_all[i]));
        }
    printf("Frozen: %zu clumps changed family, %zu of the first batch among them, their max |v| = %.3e\n", frozen, in_family, vmax_frozen);
    std::cout << "DEMdemo_FillInBatches exiting..." << std::endl;
    return 0;
}
