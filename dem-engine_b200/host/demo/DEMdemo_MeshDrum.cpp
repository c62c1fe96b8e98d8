// Clumps tumbling in a rotating drum made of triangles (the set-up family of the reference's DEMdemo_Mixer.cpp /
// DEMdemo_RotatingDrum.cpp, with the drum as a mesh so that the sphere--triangle path carries all wall contacts).
// The drum is written as a Wavefront .obj by this script and loaded back through AddWavefrontMeshObject.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>
#include <DEM/utils/Samplers.hpp>

#include <cmath>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <iostream>

using namespace deme;

// closed cylinder about the y axis, inward-facing facets
static void write_drum_obj(const std::string& path, float R, float L, int nc, int na) {
    std::ofstream f(path);
    const float PI = 3.14159265358979f;
    for (int a = 0; a <= na; a++)
        for (int c = 0; c < nc; c++) {
            const float th = 2.f * PI * c / nc;
            f << "v " << R * std::cos(th) << " " << (-L / 2 + L * a / na) << " " << R * std::sin(th) << "\n";
        }
    f << "v 0 " << -L / 2 << " 0\nv 0 " << L / 2 << " 0\n";
    auto id = [&](int a, int c) { return a * nc + (c % nc) + 1; };
    const int c0 = (na + 1) * nc + 1, c1 = c0 + 1;
    for (int a = 0; a < na; a++)
        for (int c = 0; c < nc; c++)  // one quad of the mantle as a polygon: the loader fans it into two facets
            f << "f " << id(a, c) << " " << id(a, c + 1) << " " << id(a + 1, c + 1) << " " << id(a + 1, c) << "\n";
    for (int c = 0; c < nc; c++) {
        f << "f " << c0 << " " << id(0, c + 1) << " " << id(0, c) << "\n";      // cap at -L/2, normal +y
        f << "f " << c1 << " " << id(na, c) << " " << id(na, c + 1) << "\n";    // cap at +L/2, normal -y
    }
}

int main(int argc, char** argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 5;
    DEMSolver DEMSim;
    DEMSim.SetVerbosity(QUIET);
    DEMSim.SetOutputContent(ABSV);
    DEMSim.SetMeshOutputFormat(MESH_FORMAT::VTK);

    auto mat_type_granular = DEMSim.LoadMaterial({{"E", 1e8}, {"nu", 0.3}, {"CoR", 0.5}, {"mu", 0.4}, {"Crr", 0.0}});
    auto mat_type_drum = DEMSim.LoadMaterial({{"E", 2e8}, {"nu", 0.3}, {"CoR", 0.5}, {"mu", 0.6}, {"Crr", 0.0}});
    DEMSim.SetMaterialPropertyPair("mu", mat_type_granular, mat_type_drum, 0.6);

    const float R = 0.1f, L = 0.08f;
    DEMSim.InstructBoxDomainDimension(0.3, 0.3, 0.3);

    std::filesystem::path out_dir = std::filesystem::current_path() / "DemoOutput_MeshDrum";
    std::filesystem::create_directories(out_dir);
    const std::string obj = (out_dir / "drum.obj").string();
    write_drum_obj(obj, R, L, 96, 8);
    auto drum = DEMSim.AddWavefrontMeshObject(obj, mat_type_drum);
    std::cout << "Drum mesh: " << drum->GetNumTriangles() << " facets, " << drum->GetNumNodes() << " nodes" << std::endl;
    drum->SetFamily(10);
    drum->SetMass(1.f);
    drum->SetMOI(make_float3(1.f, 1.f, 1.f));
    DEMSim.SetFamilyPrescribedAngVel(10, "0", "6.0", "0");
    DEMSim.SetFamilyPrescribedLinVel(10, "0", "0", "0");
    auto drum_tracker = DEMSim.Track(drum);

    // three-sphere clumps filling the lower part of the drum
    const float scale = 0.003f;
    DEMClumpTemplate shape;
    shape.ReadComponentFromFile(GetDEMEDataFile("clumps/3_clump.csv"));
    shape.Scale(scale);
    shape.SetMass(2.6e3f * 5.5886717f * scale * scale * scale);
    const float moi_s = 2.6e3f * scale * scale * scale * scale * scale;
    shape.SetMOI(make_float3(2.928f, 2.6029f, 3.9908f) * moi_s);
    shape.SetMaterial(mat_type_granular);
    auto clump_type = DEMSim.LoadClumpType(shape);

    HCPSampler sampler(scale * 3.2f);
    std::vector<float3> all = sampler.SampleCylinderY(make_float3(0, 0, 0), R - 4 * scale, L / 2 - 3 * scale), xyz;
    for (const auto& p : all)
        if (p.z < -0.02f) xyz.push_back(p);
    auto particles = DEMSim.AddClumps(clump_type, xyz);
    particles->SetFamily(0);
    std::cout << xyz.size() << " clumps" << std::endl;

    DEMSim.SetInitTimeStep(5e-6);
    DEMSim.SetGravitationalAcceleration(make_float3(0, 0, -9.81));
    DEMSim.SetCDUpdateFreq(20);
    DEMSim.SetMaxVelocity(5.);
    DEMSim.SetExpandSafetyAdder(1.0);
    DEMSim.SetErrorOutVelocity(50.);
    DEMSim.Initialize();

    auto max_v = DEMSim.CreateInspector("clump_max_absv");
    auto tracker = DEMSim.Track(particles);
    for (int i = 0; i < frames; i++) {
        DEMSim.DoDynamicsThenSync(0.02);
        float rmax = 0.f, ymax = 0.f;
        for (const auto& p : tracker->Positions()) {
            rmax = std::max(rmax, std::sqrt(p.x * p.x + p.z * p.z));
            ymax = std::max(ymax, std::fabs(p.y));
        }
        const float4 q = drum_tracker->OriQ();
        const float angle = 2.f * std::atan2(q.y, q.w);
        printf("Frame %d: t = %.4f, max radial = %.5f, max |y| = %.5f, drum angle = %.5f, max v = %.4f, contacts = %zu\n", i,
               DEMSim.GetSimTime(), rmax, ymax, angle, max_v->GetValue(), DEMSim.GetNumContacts());
        char name[64];
        snprintf(name, sizeof(name), "drum_%04d.vtk", i);
        DEMSim.WriteMeshFile(out_dir / name);
    }
    DEMSim.ShowTimingStats();
    std::cout << "DEMdemo_MeshDrum exiting..." << std::endl;
    return 0;
}
