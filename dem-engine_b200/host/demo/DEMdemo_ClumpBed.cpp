// A demo script written against the reference's public API (the same calls DEMdemo_Mixer.cpp / DEMdemo_BallDrop.cpp
// make): three-sphere clumps poured into a box with a cylindrical wall, a plane pushed down by a prescribed
// velocity, a tracker and inspectors.  It compiles unchanged against either library's <DEM/API.h>.
#include <core/ApiVersion.h>
#include <core/utils/ThreadManager.h>
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <chrono>
#include <cstdio>
#include <filesystem>

using namespace deme;
using namespace std::filesystem;

int main(int argc, char** argv) {
    const int n_frames = argc > 1 ? atoi(argv[1]) : 5;
    DEMSolver DEMSim;
    DEMSim.SetVerbosity(INFO);
    DEMSim.SetOutputFormat(OUTPUT_FORMAT::CSV);
    DEMSim.SetOutputContent(OUTPUT_CONTENT::ABSV | OUTPUT_CONTENT::VEL | OUTPUT_CONTENT::FAMILY);
    DEMSim.SetNoForceRecord();

    auto mat_type_walls = DEMSim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.6}, {"mu", 0.5}, {"Crr", 0.0}});
    auto mat_type_granular = DEMSim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.6}, {"mu", 0.2}, {"Crr", 0.0}});
    DEMSim.SetMaterialPropertyPair("mu", mat_type_walls, mat_type_granular, 0.5);

    const float step_size = 5e-6;
    const double world_size = 0.3;
    DEMSim.InstructBoxDomainDimension(world_size, world_size, world_size);
    DEMSim.InstructBoxDomainBoundingBC("all", mat_type_granular);

    auto walls = DEMSim.AddExternalObject();
    walls->AddCylinder(make_float3(0), make_float3(0, 0, 1), world_size / 2., mat_type_walls, 0);

    const float granular_rad = 0.005;
    float mass = 2.6e3 * 5.5886717;
    float3 MOI = make_float3(2.928, 2.6029, 3.9908) * 2.6e3;
    std::shared_ptr<DEMClumpTemplate> template_granular =
        DEMSim.LoadClumpType(mass, MOI, GetDEMEDataFile("clumps/3_clump.csv").string(), mat_type_granular);
    template_granular->Scale(granular_rad);

    HCPSampler sampler(3.f * granular_rad);
    const float fill_height = world_size / 3.;
    float3 fill_center = make_float3(0, 0, -world_size / 2. + fill_height / 2. + 2 * granular_rad);
    const float fill_radius = world_size / 2. - 2. * granular_rad;
    auto input_xyz = sampler.SampleCylinderZ(fill_center, fill_radius, fill_height / 2);
    auto particles = DEMSim.AddClumps(template_granular, input_xyz);
    particles->SetVel(make_float3(0, 0, -0.5));
    std::cout << "Total num of particles: " << input_xyz.size() << std::endl;

    // a lid that moves down at constant speed
    auto lid = DEMSim.AddExternalObject();
    lid->AddPlane(make_float3(0, 0, 0), make_float3(0, 0, -1), mat_type_walls);
    lid->SetInitPos(make_float3(0, 0, 0.0));
    lid->SetFamily(10);
    DEMSim.SetFamilyPrescribedLinVel(10, "0", "0", "-0.2");
    DEMSim.SetFamilyFixed(11);
    auto lid_tracker = DEMSim.Track(lid);
    auto particle_tracker = DEMSim.Track(particles);

    DEMSim.SetInitTimeStep(step_size);
    DEMSim.SetGravitationalAcceleration(make_float3(0, 0, -9.81));
    DEMSim.SetCDUpdateFreq(20);
    DEMSim.SetExpandSafetyAdder(2.0);
    DEMSim.SetErrorOutVelocity(20.);
    DEMSim.Initialize();

    auto max_z_finder = DEMSim.CreateInspector("clump_max_z");
    auto max_v_finder = DEMSim.CreateInspector("clump_max_absv");
    auto ke_finder = DEMSim.CreateInspector("clump_kinetic_energy");

    path out_dir = current_path() / "DemoOutput_ClumpBed";
    create_directories(out_dir);
    const float frame_time = 0.005;
    auto start = std::chrono::high_resolution_clock::now();
    for (int frame = 0; frame < n_frames; frame++) {
        char filename[100];
        sprintf(filename, "DEMdemo_output_%04d.csv", frame);
        DEMSim.WriteClumpFile(out_dir / filename);
        DEMSim.DoDynamics(frame_time);
        float3 lid_pos = lid_tracker->Pos();
        float3 p0 = particle_tracker->Pos(0);
        std::cout << "Frame " << frame << ": t = " << DEMSim.GetSimTime() << ", lid z = " << lid_pos.z
                  << ", max z = " << max_z_finder->GetValue() << ", max v = " << max_v_finder->GetValue()
                  << ", KE = " << ke_finder->GetValue() << ", contacts = " << DEMSim.GetNumContacts()
                  << ", clump 0 at (" << p0.x << ", " << p0.y << ", " << p0.z << ")" << std::endl;
    }
    // stop the lid: on-the-fly family change, as the reference's ChangeFamily
    DEMSim.ChangeFamily(10, 11);
    DEMSim.DoDynamicsThenSync(frame_time);
    std::cout << "Lid after being fixed: z = " << lid_tracker->Pos().z << " v_z = " << lid_tracker->Vel().z << std::endl;
    std::chrono::duration<double> time_sec = std::chrono::high_resolution_clock::now() - start;
    std::cout << time_sec.count() << " seconds (wall time) to finish the simulation" << std::endl;
    DEMSim.ShowThreadCollaborationStats();
    DEMSim.ShowTimingStats();
    DEMSim.ShowMemStats();
    std::cout << "DEMdemo_ClumpBed exiting..." << std::endl;
    return 0;
}
