// Checkpoint / restart through the on-disk formats either side of the stepping path: a settling clump bed writes its
// clump file and contact file (with the Hertz-Mindlin history wildcards); a second solver is built from those files
// alone (ReadClump*FromCsv, ReadContactPairsFromCsv, ReadContactWildcardsFromCsv, SetExistingContacts /
// SetExistingContactWildcards -- the restart recipe of the reference, DEMdemo_GRCPrep_Part2.cpp:60-110) and both run on.
// Prints how far the restarted bed is from the uninterrupted one after the same number of further steps.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <cmath>
#include <cstdio>
#include <filesystem>
#include <iostream>

using namespace deme;

static std::shared_ptr<DEMClumpTemplate> setup(DEMSolver& sim, std::shared_ptr<DEMMaterial>& mat) {
    sim.SetVerbosity(QUIET);
    sim.SetOutputContent({"XYZ", "QUAT", "VEL", "ANG_VEL", "FAMILY"});
    // what a restart needs from the contact file: the geometry ids of each pair and its history words
    sim.SetContactOutputContent({"OWNER", "GEO_ID", "FORCE", "CNT_WILDCARD"});
    mat = sim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.5}, {"mu", 0.4}, {"Crr", 0.0}});
    auto tmpl = sim.LoadClumpType(2.6e3f * 5.5886717f, make_float3(2.928f, 2.6029f, 3.9908f) * 2.6e3f,
                                  GetDEMEDataFile("clumps/3_clump.csv").string(), mat);
    tmpl->Scale(0.01f);
    tmpl->AssignName("three_sphere");
    sim.InstructBoxDomainDimension(0.4, 0.4, 0.6);
    sim.InstructBoxDomainBoundingBC("top_open", mat);
    sim.SetGravitationalAcceleration(make_float3(0, 0, -9.81));
    sim.SetInitTimeStep(1e-5);
    sim.SetCDUpdateFreq(10);
    sim.UseAdaptiveUpdateFreq(false);  // (this script compares contact lists: keep the cycle length, hence the margin, fixed)
    return tmpl;
}

int main() {
    namespace fs = std::filesystem;
    const fs::path dir = fs::current_path() / "DemoOutput_Restart";
    fs::create_directories(dir);
    std::shared_ptr<DEMMaterial> mat1, mat2;
    DEMSolver A;
    auto tA = setup(A, mat1);
    HCPSampler sampler(0.03f);
    auto xyz = sampler.SampleBox(make_float3(0, 0, -0.2f), make_float3(0.15f, 0.15f, 0.06f));
    auto pA = A.AddClumps(tA, xyz);
    pA->SetVel(make_float3(0.05f, 0, -0.5f));
    auto trA = A.Track(pA);
    A.Initialize();
    A.DoDynamicsThenSync(0.08);   // the bed lands and is full of frictional contacts
    A.WriteClumpFile(dir / "clumps.csv", 9);
    A.WriteContactFile(dir / "contacts.csv");
    const size_t n = pA->GetNumClumps();

    // ---- the restarted solver, from the two files only ----
    DEMSolver B;
    auto tB = setup(B, mat2);
    auto pos = DEMSolver::ReadClumpXyzFromCsv((dir / "clumps.csv").string());
    auto quat = DEMSolver::ReadClumpQuatFromCsv((dir / "clumps.csv").string());
    auto vel = DEMSolver::ReadClumpVelFromCsv((dir / "clumps.csv").string());
    auto angvel = DEMSolver::ReadClumpAngVelFromCsv((dir / "clumps.csv").string());
    auto pairs = DEMSolver::ReadContactPairsFromCsv((dir / "contacts.csv").string());
    auto wildcards = DEMSolver::ReadContactWildcardsFromCsv((dir / "contacts.csv").string());
    auto pB = B.AddClumps(tB, pos["three_sphere"]);
    pB->SetOriQ(quat["three_sphere"]);
    pB->SetVel(vel["three_sphere"]);
    pB->SetAngVel(angvel["three_sphere"]);
    pB->SetExistingContacts(pairs);
    pB->SetExistingContactWildcards(wildcards);
    auto trB = B.Track(pB);
    B.Initialize();
    // what the restarted solver now holds (every listed pair): the test compares it with contacts.csv entry by entry
    B.WriteContactFile(dir / "contacts_restarted.csv", -1.0f);

    // ---- a third one restarted WITHOUT the contact history, for scale ----
    std::shared_ptr<DEMMaterial> mat3;
    DEMSolver C;
    auto tC = setup(C, mat3);
    auto pC = C.AddClumps(tC, pos["three_sphere"]);
    pC->SetOriQ(quat["three_sphere"]);
    pC->SetVel(vel["three_sphere"]);
    pC->SetAngVel(angvel["three_sphere"]);
    auto trC = C.Track(pC);
    C.Initialize();

    const double more = 0.005;
    A.DoDynamicsThenSync(more);
    B.DoDynamicsThenSync(more);
    C.DoDynamicsThenSync(more);
    auto xa = trA->Positions(), xb = trB->Positions(), xc = trC->Positions();
    auto va = trA->Velocities(), vb = trB->Velocities(), vc = trC->Velocities();
    double dxb = 0, dxc = 0, dvb = 0, dvc = 0;
    for (size_t i = 0; i < n; i++) {
        dxb = std::max<double>(dxb, length(xb[i] - xa[i]));
        dxc = std::max<double>(dxc, length(xc[i] - xa[i]));
        dvb = std::max<double>(dvb, length(vb[i] - va[i]));
        dvc = std::max<double>(dvc, length(vc[i] - va[i]));
    }
    printf("clumps = %zu, contact pairs read = %zu, wildcard columns = %zu\n", n, pairs.size(), wildcards.size());
    printf("restart with history   : max |dx| = %.3e, max |dv| = %.3e\n", dxb, dvb);
    printf("restart without history: max |dx| = %.3e, max |dv| = %.3e\n", dxc, dvc);
    std::cout << "DEMdemo_Restart exiting..." << std::endl;
    return 0;
}
