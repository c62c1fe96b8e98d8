// A plate shaken with a time-dependent prescribed velocity, grains bouncing on it (the pattern of the reference's
// DEMdemo_Sieve.cpp:95-101 and DEMdemo_Shake.cpp: SetFamilyPrescribedLinVel(fam, "0", "0", "<expression of t>")).
// Prints the plate's position against the sum the explicit integrator must produce, z_n = sum_k v(t_k) h.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <cmath>
#include <cstdio>
#include <iostream>

using namespace deme;

int main() {
    DEMSolver DEMSim;
    DEMSim.SetVerbosity(QUIET);
    auto mat = DEMSim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.6}, {"mu", 0.3}, {"Crr", 0.0}});
    auto ball = DEMSim.LoadSphereType(2600.f * 4.f / 3.f * 3.14159f * 0.005f * 0.005f * 0.005f, 0.005f, mat);
    DEMSim.InstructBoxDomainDimension(0.3, 0.3, 0.4);
    DEMSim.InstructBoxDomainBoundingBC("only_sides", mat);
    DEMSim.SetGravitationalAcceleration(make_float3(0, 0, -9.81));
    const double h = 1e-5;
    DEMSim.SetInitTimeStep(h);
    DEMSim.SetCDUpdateFreq(10);

    auto plate = DEMSim.AddExternalObject();
    plate->AddPlane(make_float3(0, 0, 0), make_float3(0, 0, 1), mat);
    plate->SetInitPos(make_float3(0, 0, -0.1f));
    plate->SetFamily(1);
    // z: a 50 Hz shake; x: a constant drift that only starts at t = 4 ms; constant arithmetic in y
    DEMSim.SetFamilyPrescribedLinVel(1, "(t > 0.004) ? 0.5 : 0", "-0.1 / 4", "0.3 * sin(2 * deme::PI * 50 * t)");
    DEMSim.SetFamilyPrescribedAngVel(1, "0", "0", "0");
    auto plate_tracker = DEMSim.Track(plate);

    GridSampler grid(0.0125f);
    auto grains = DEMSim.AddClumps(ball, grid.SampleBox(make_float3(0, 0, -0.07f), make_float3(0.1f, 0.1f, 0.02f)));
    auto max_z = DEMSim.CreateInspector("clump_max_z");
    DEMSim.Initialize();

    double zsum = -0.1, xsum = 0.0, ysum = 0.0, t = 0.0;
    for (int frame = 1; frame <= 4; frame++) {
        DEMSim.DoDynamicsThenSync(250 * h);
        // (the reference's step loop runs on the float-rounded h, so a call may take one step more than duration / h:
        // follow the solver's own clock)
        for (; t < DEMSim.GetSimTime() - 0.5 * h; t += (double)(float)h) {
            zsum += (double)(float)(0.3 * std::sin(2 * 3.14159265358979323846 * 50 * t)) * (double)(float)h;
            xsum += ((t > 0.004) ? 0.5 : 0.0) * (double)(float)h;
            ysum += (double)(float)(-0.1 / 4) * (double)(float)h;
        }
        const float3 p = plate_tracker->Pos();
        printf("Frame %d: t = %.5f, plate = (%.7f, %.7f, %.7f), expected = (%.7f, %.7f, %.7f), grains max z = %.5f\n", frame,
               DEMSim.GetSimTime(), p.x, p.y, p.z, xsum, ysum, zsum, max_z->GetValue());
    }
    printf("check: sin(3) via the parser = %.9f\n", DEMSolver::EvaluatePrescription("sin(t)", 3.0));
    std::cout << "DEMdemo_Shaker exiting..." << std::endl;
    return 0;
}
