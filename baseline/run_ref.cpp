// run_ref.cpp -- drives the UNMODIFIED reference (projectchrono/DEM-Engine, built by baseline/build_ref.sh into
// baseline/_ref/build) through its own public API (src/DEM/API.h of the reference) on the scene files bench.py /
// tools/run_reference_gpu.py write.  Built twice from this ONE source: baseline/_ref/run_ref links the unmodified reference
// (baseline/build_ref.sh; nothing of this repository's engine is linked or called there), dem-engine_b200/host/run_b200
// links this repository's deme::DEMSolver facade instead -- the same script drives both, which is the drop-in claim.
//
//   run_ref bench  <scene.bin> <nGPUs> <steps> <warmup_steps> [cd_update_freq (0 = the reference's adaptive default)]
//       builds the bed, Initialize(), DoDynamicsThenSync(warmup*h), then times DoDynamicsThenSync(steps*h)
//       (src/DEM/APIPublic.cpp:2446-2479) with a wall clock, as SURVEY.md 8(d) prescribes; prints one JSON line.
//   run_ref e2e    <scene.bin> <nGPUs> <frames> <warmup_steps> [cd_update_freq]
//       every frame: DoDynamics(h) (one step), then a tracked clump's Pos / Vel and two inspectors read back to the host
//   run_ref parity <scene.bin> <steps_per_checkpoint> <n_checkpoints> <out.bin>
//       the lock-step recipe of DEMdemo_TestPack.cpp:30-45 (SetCDUpdateFreq(0), UseAdaptiveUpdateFreq(false),
//       DisableAdaptiveBinSize, UseCubForceCollection, jitify of templates off); dumps pos / quat / vel / angvel of
//       every clump after each checkpoint (float32, the finest the reference's getters report).
//
// Scene file (little endian): char[4] "DEMS", u32 version = 1, u32 nClumps, f32 box[3], f32 scale, f32 h,
//   f32 E, nu, CoR, mu, Crr, f32 beta (<0: velocity based margin), u32 has_vel,
//   f32 xyz[n][3], f32 quat_wxyz[n][4], then if has_vel: f32 vel[n][3], f32 omgBar[n][3]
// The clump is data/clumps/3_clump.csv scaled by `scale` with the Mixer demo's mass / MOI (DEMdemo_Mixer.cpp:68-72).
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// which engine this binary was linked with (the facade's Makefile passes -DRUN_IMPL_NAME=...)
#ifndef RUN_IMPL_NAME
    #define RUN_IMPL_NAME "DEME (unmodified reference)"
#endif

using namespace deme;

struct SceneFile {
    uint32_t n = 0, has_vel = 0;
    float box[3], scale, h, E, nu, CoR, mu, Crr, beta;
    std::vector<float3> xyz, vel, omg;
    std::vector<float4> quat;  // (x, y, z, w) as the reference stores it (src/DEM/dT.cpp:3024)
};

static bool read_scene(const char* path, SceneFile& s) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char magic[4];
    uint32_t version = 0;
    bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "DEMS", 4) == 0 && fread(&version, 4, 1, f) == 1 && version == 1;
    ok = ok && fread(&s.n, 4, 1, f) == 1 && fread(s.box, 4, 3, f) == 3 && fread(&s.scale, 4, 1, f) == 1 &&
         fread(&s.h, 4, 1, f) == 1 && fread(&s.E, 4, 1, f) == 1 && fread(&s.nu, 4, 1, f) == 1 &&
         fread(&s.CoR, 4, 1, f) == 1 && fread(&s.mu, 4, 1, f) == 1 && fread(&s.Crr, 4, 1, f) == 1 &&
         fread(&s.beta, 4, 1, f) == 1 && fread(&s.has_vel, 4, 1, f) == 1;
    if (ok) {
        std::vector<float> buf((size_t)s.n * 4);
        s.xyz.resize(s.n);
        s.quat.resize(s.n);
        ok = fread(buf.data(), 4, (size_t)s.n * 3, f) == (size_t)s.n * 3;
        for (uint32_t i = 0; ok && i < s.n; i++) s.xyz[i] = make_float3(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
        ok = ok && fread(buf.data(), 4, (size_t)s.n * 4, f) == (size_t)s.n * 4;
        for (uint32_t i = 0; ok && i < s.n; i++) s.quat[i] = make_float4(buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3], buf[4 * i]);
        if (ok && s.has_vel) {
            s.vel.resize(s.n);
            s.omg.resize(s.n);
            ok = fread(buf.data(), 4, (size_t)s.n * 3, f) == (size_t)s.n * 3;
            for (uint32_t i = 0; ok && i < s.n; i++) s.vel[i] = make_float3(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
            ok = ok && fread(buf.data(), 4, (size_t)s.n * 3, f) == (size_t)s.n * 3;
            for (uint32_t i = 0; ok && i < s.n; i++) s.omg[i] = make_float3(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
        }
    }
    fclose(f);
    return ok;
}

static std::shared_ptr<DEMClumpBatch> build(DEMSolver& sim, const SceneFile& s) {
    sim.SetVerbosity(ERR);
    sim.SetOutputFormat(OUTPUT_FORMAT::CSV);
    auto mat = sim.LoadMaterial({{"E", s.E}, {"nu", s.nu}, {"CoR", s.CoR}, {"mu", s.mu}, {"Crr", s.Crr}});
    const float mass = 2.6e3f * 5.5886717f;
    const float3 MOI = make_float3(2.928f, 2.6029f, 3.9908f) * 2.6e3f;
    auto tmpl = sim.LoadClumpType(mass, MOI, GetDEMEDataFile("clumps/3_clump.csv"), mat);
    tmpl->Scale(s.scale);
    sim.InstructBoxDomainDimension(s.box[0], s.box[1], s.box[2]);
    sim.InstructBoxDomainBoundingBC("top_open", mat);
    auto batch = sim.AddClumps(tmpl, s.xyz);
    batch->SetOriQ(s.quat);
    if (s.has_vel) {
        batch->SetVel(s.vel);
        batch->SetAngVel(s.omg);
    }
    sim.SetInitTimeStep(s.h);
    sim.SetGravitationalAcceleration(make_float3(0, 0, -9.81f));
    return batch;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: run_ref bench|parity ...\n");
        return 2;
    }
    const std::string mode = argv[1];
    SceneFile s;
    if (argc < 3 || !read_scene(argv[2], s)) {
        fprintf(stderr, "run_ref: cannot read scene file\n");
        return 2;
    }
    try {
        if (mode == "bench") {
            if (argc < 6) return 2;
            const int ngpu = atoi(argv[3]);
            const long steps = atol(argv[4]), warm = atol(argv[5]);
            const int freq = argc > 6 ? atoi(argv[6]) : 0;
            const auto t_init0 = std::chrono::steady_clock::now();
            DEMSolver sim(ngpu);
            build(sim, s);
            if (freq > 0) {
                sim.SetCDUpdateFreq(freq);
                sim.UseAdaptiveUpdateFreq(false);
            }
            sim.Initialize();
            const double init_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init0).count();
            sim.DoDynamicsThenSync((double)warm * (double)s.h);
            const auto t0 = std::chrono::steady_clock::now();
            sim.DoDynamicsThenSync((double)steps * (double)s.h);
            const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("{\"impl\": \"" RUN_IMPL_NAME "\", \"n_gpus\": %d, \"clumps\": %u, \"steps\": %ld, \"warmup\": %ld, "
                   "\"wall_s\": %.6f, \"steps_per_s\": %.3f, \"init_s\": %.2f, \"n_contacts\": %zu, \"update_freq\": %.2f, "
                   "\"cd_update_freq_setting\": %d}\n",
                   ngpu, s.n, steps, warm, wall, (double)steps / wall, init_s, sim.GetNumContacts(), sim.GetUpdateFreq(), freq);
            fflush(stdout);
            sim.ShowThreadCollaborationStats();
            sim.ShowTimingStats();
        } else if (mode == "e2e") {
            // what a co-simulating caller does every step: advance one step, read a tracked body and two inspectors
            // back to the host (DEMdemo_Mixer.cpp:130-140 polls its inspectors the same way)
            if (argc < 6) return 2;
            const int ngpu = atoi(argv[3]);
            const long frames = atol(argv[4]), warm = atol(argv[5]);
            const int freq = argc > 6 ? atoi(argv[6]) : 0;
            DEMSolver sim(ngpu);
            auto batch = build(sim, s);
            if (freq > 0) {
                sim.SetCDUpdateFreq(freq);
                sim.UseAdaptiveUpdateFreq(false);
            }
            auto tracker = sim.Track(batch);
            auto max_v = sim.CreateInspector("clump_max_absv");
            auto ke = sim.CreateInspector("clump_kinetic_energy");
            sim.Initialize();
            sim.DoDynamicsThenSync((double)warm * (double)s.h);
            const size_t probe = s.n / 2;
            double checksum = 0.0;
            const auto t0 = std::chrono::steady_clock::now();
            for (long k = 0; k < frames; k++) {
                sim.DoDynamics((double)s.h);
                const float3 p = tracker->Pos(probe);
                const float3 v = tracker->Vel(probe);
                checksum += (double)p.z + (double)v.z + (double)max_v->GetValue() + (double)ke->GetValue();
            }
            const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("{\"impl\": \"deme::DEMSolver API\", \"mode\": \"e2e\", \"n_gpus\": %d, \"clumps\": %u, \"frames\": %ld, \"warmup\": %ld, "
                   "\"wall_s\": %.6f, \"steps_per_s\": %.3f, \"d2h_bytes_per_step\": 32, \"checksum\": %.9g}\n",
                   ngpu, s.n, frames, warm, wall, (double)frames / wall, checksum);
        } else if (mode == "parity") {
            if (argc < 6) return 2;
            const long per = atol(argv[3]);
            const int ncp = atoi(argv[4]);
            DEMSolver sim(1);
            build(sim, s);
            sim.SetCDUpdateFreq(0);
            sim.UseAdaptiveUpdateFreq(false);
            sim.DisableAdaptiveBinSize();
            sim.DisableJitifyClumpTemplates();
            sim.DisableJitifyMassProperties();
            sim.UseCubForceCollection();
            if (s.beta >= 0.f) sim.SetExpandFactor(s.beta, true);
            sim.Initialize();
            FILE* out = fopen(argv[5], "wb");
            if (!out) return 2;
            const uint32_t hdr[4] = {0x50464544u /* "DEFP" */, s.n, (uint32_t)ncp, (uint32_t)per};
            fwrite(hdr, 4, 4, out);
            for (int c = 0; c < ncp; c++) {
                for (long k = 0; k < per; k++) sim.DoStepDynamics();  // each is DoDynamics(h): exactly one step (API.h:1263)
                const auto P = sim.GetOwnerPosition(0, s.n);
                const auto Q = sim.GetOwnerOriQ(0, s.n);
                const auto V = sim.GetOwnerVelocity(0, s.n);
                const auto W = sim.GetOwnerAngVel(0, s.n);
                fwrite(P.data(), sizeof(float3), s.n, out);
                fwrite(Q.data(), sizeof(float4), s.n, out);  // x, y, z, w
                fwrite(V.data(), sizeof(float3), s.n, out);
                fwrite(W.data(), sizeof(float3), s.n, out);
            }
            fclose(out);
            printf("{\"impl\": \"" RUN_IMPL_NAME "\", \"mode\": \"parity\", \"clumps\": %u, \"checkpoints\": %d, "
                   "\"steps_per_checkpoint\": %ld, \"n_contacts\": %zu}\n", s.n, ncp, per, sim.GetNumContacts());
        } else {
            return 2;
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "run_ref: %s\n", e.what());
        return 1;
    }
    return 0;
}
