// Two spheres collide head on (the set-up of the reference's install check DEMdemo_SingleSphereCollide.cpp:49-63):
// prints the measured coefficient of restitution, which the Hertzian model with CoR = 0.6 must reproduce.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <cstdio>
#include <iostream>

using namespace deme;

int main() {
    DEMSolver DEMSim;
    DEMSim.SetVerbosity(QUIET);
    auto mat_type_1 = DEMSim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.6}, {"mu", 0.0}, {"Crr", 0.0}});
    auto sph_type_1 = DEMSim.LoadSphereType(11728., 1., mat_type_1);
    std::vector<float3> input_xyz1(1, make_float3(-1.05, 0, 0)), input_xyz2(1, make_float3(1.05, 0, 0));
    auto particles1 = DEMSim.AddClumps(sph_type_1, input_xyz1);
    particles1->SetVel(make_float3(1.f, 0, 0));
    particles1->SetFamily(0);
    auto tracker1 = DEMSim.Track(particles1);
    auto particles2 = DEMSim.AddClumps(sph_type_1, input_xyz2);
    particles2->SetVel(make_float3(-1.f, 0, 0));
    particles2->SetFamily(1);
    auto tracker2 = DEMSim.Track(particles2);
    DEMSim.InstructBoxDomainDimension(10, 10, 10);
    DEMSim.SetGravitationalAcceleration(make_float3(0, 0, 0));
    DEMSim.SetInitTimeStep(2e-5);
    DEMSim.SetCDUpdateFreq(10);
    DEMSim.SetMaxVelocity(3.);
    DEMSim.SetExpandSafetyAdder(1.0);
    DEMSim.Initialize();
    for (int i = 0; i < 100; i++) DEMSim.DoDynamics(1e-3);
    const float v1 = tracker1->Vel().x, v2 = tracker2->Vel().x;
    printf("v1 after = %.6f, v2 after = %.6f, CoR measured = %.4f\n", v1, v2, (v2 - v1) / 2.0);
    printf("x1 = %.6f x2 = %.6f contacts = %zu\n", tracker1->Pos().x, tracker2->Pos().x, DEMSim.GetNumContacts());
    return 0;
}
