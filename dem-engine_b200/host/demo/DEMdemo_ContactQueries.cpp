// Contact read-outs and the small per-owner controls of the facade on a scene small enough to reason about: a 4 x 4 x 2
// block of spheres settled on the floor of a box, plus one sphere far above everything.  Every check prints PASS / FAIL;
// the exit code is the number of failed checks.
//   GetContactDetailedInfo (fields, normals of floor contacts, total floor force = weight), SetContactWildcardValue /
//   SetFamilyContactWildcardValue, MarkPersistentContact / Remove..., region inspectors, AddAcc (next step only),
//   UpdateSimParams.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>

#include <cmath>
#include <cstdio>
#include <string>

using namespace deme;

static int n_failed = 0;
static void report(bool ok, const std::string& what) {
    printf("%s  %s\n", ok ? "PASS" : "FAIL", what.c_str());
    if (!ok) n_failed++;
}

int main() {
    DEMSolver DEMSim;
    DEMSim.SetVerbosity(QUIET);
    DEMSim.SetContactOutputContent(OWNER | FORCE | CNT_POINT | NORMAL | CNT_WILDCARD | GEO_ID);
    auto mat = DEMSim.LoadMaterial({{"E", 1e7}, {"nu", 0.3}, {"CoR", 0.3}, {"mu", 0.4}, {"Crr", 0.0}});
    const float R = 0.01f, mass = 2600.f * 4.f / 3.f * (float)PI * R * R * R;
    auto ball = DEMSim.LoadSphereType(mass, R, mat);
    DEMSim.InstructBoxDomainDimension({-0.1, 0.1}, {-0.1, 0.1}, {0.0, 1.0});
    DEMSim.InstructBoxDomainBoundingBC("top_open", mat);

    std::vector<float3> xyz;
    std::vector<unsigned int> fam;
    for (int k = 0; k < 2; k++)
        for (int j = 0; j < 4; j++)
            for (int i = 0; i < 4; i++) {
                xyz.push_back(make_float3(-0.03f + 0.02f * i, -0.03f + 0.02f * j, R + 0.0199f * k));
                fam.push_back(k);  // bottom layer: family 0, top layer: family 1
            }
    auto bed = DEMSim.AddClumps(ball, xyz);
    bed->SetFamilies(fam);
    auto bed_tracker = DEMSim.Track(bed);
    auto loner = DEMSim.AddClumps(ball, make_float3(0.05f, 0.05f, 0.8f));
    loner->SetFamily(2);
    auto loner_tracker = DEMSim.Track(loner);

    auto max_z_all = DEMSim.CreateInspector("clump_max_z");
    auto max_z_low = DEMSim.CreateInspector("clump_max_z", "return Z < 0.5;");
    auto mass_left = DEMSim.CreateInspector("clump_mass", "return X < 0;");
    auto volume = DEMSim.CreateInspector("clump_volume");
    auto min_z_all = DEMSim.CreateInspector("clump_min_z");
    auto max_v_all = DEMSim.CreateInspector("clump_max_absv");
    auto max_v_low = DEMSim.CreateInspector("clump_max_absv", "return Z < 0.5;");

    // 2^-18 s: exactly representable as a float, so that DoStepDynamics() is ONE step.  (DoDynamics(t) steps while
    // "cycle < t" with cycle advanced by (double)(float)h, dT.cpp:2401 of the reference: a step size whose float value is
    // below its double value -- 5e-6, say -- makes DoStepDynamics() take two steps there, and here.)
    const double h = 1.0 / 262144.0;
    DEMSim.SetInitTimeStep(h);
    DEMSim.SetGravitationalAcceleration(make_float3(0, 0, -9.81f));
    DEMSim.SetCDUpdateFreq(10);
    DEMSim.UseAdaptiveUpdateFreq(false);
    DEMSim.SetExpandSafetyAdder(0.5);
    DEMSim.Initialize();

    DEMSim.DoDynamicsThenSync(0.05);  // the block settles (it starts nearly at rest)

    // ---- inspectors, whole domain and regions ----
    {
        const float zl = loner_tracker->Pos().z;
        report(std::fabs(max_z_all->GetValue() - (zl + R)) < 1e-5f, "clump_max_z over the domain is the top of the lone sphere");
        float top = -1e30f;
        double m_left = 0;
        const auto pos = bed_tracker->Positions();
        for (const auto& p : pos) {
            top = std::max(top, p.z + R);
            if (p.x < 0) m_left += mass;
        }
        report(std::fabs(max_z_low->GetValue() - top) < 1e-6f, "clump_max_z with region Z < 0.5 is the top of the block");
        report(std::fabs(mass_left->GetValue() - m_left) < 1e-4 * m_left, "clump_mass with region X < 0 is half of the block");
        report(volume->GetValue() >= 0.f, "clump_volume is summed from the templates");
        float bottom = 1e30f, vmax = 0.f;
        for (const auto& p : pos) bottom = std::min(bottom, p.z - R);
        for (const auto& v : bed_tracker->Velocities()) vmax = std::max(vmax, length(v));
        report(std::fabs(min_z_all->GetValue() - bottom) < 1e-6f, "clump_min_z is the underside of the lowest sphere");
        report(std::fabs(max_v_all->GetValue() - length(loner_tracker->Vel())) < 1e-5f, "clump_max_absv is the speed of the falling sphere");
        report(std::fabs(max_v_low->GetValue() - vmax) < 1e-6f, "clump_max_absv with region Z < 0.5 is the fastest sphere of the block");
    }

    // ---- the detailed contact read-out ----
    size_t n_touching = 0;
    {
        auto info = DEMSim.GetContactDetailedInfo(1e-6f);
        n_touching = info->Size();
        double floor_force = 0;
        bool normals_ok = true, fields_ok = true;
        size_t n_floor = 0;
        for (size_t i = 0; i < info->Size(); i++) {
            const float3 n = info->GetNormal()[i];
            fields_ok = fields_ok && std::fabs(length(n) - 1.f) < 1e-4f && info->GetAOwner()[i] < 33;
            if (info->GetContactType()[i] == "SA" && std::fabs(info->GetForce()[i].z) > 0.5f * length(info->GetForce()[i]) &&
                info->GetPoint()[i].z < 0.5f * R) {
                // a sphere on the floor: body A's outward normal points down, the force on A points up
                n_floor++;
                floor_force += info->GetForce()[i].z;
                normals_ok = normals_ok && n.z < -0.999f && info->GetForce()[i].z > 0.f && info->GetAOwnerFamily()[i] == 0;
            }
        }
        report(n_touching >= 16 + 16 && fields_ok, "GetContactDetailedInfo lists the touching pairs with unit normals (" +
                                                       std::to_string(n_touching) + ")");
        report(n_floor == 16 && normals_ok, "16 floor contacts, normals (0, 0, -1), force up, family 0");
        const double weight = 32.0 * mass * 9.81;
        report(std::fabs(floor_force - weight) < 0.05 * weight,
               "the floor carries the weight of the block (" + std::to_string(floor_force) + " N of " + std::to_string(weight) + ")");
        auto all = DEMSim.GetContactDetailedInfo(-1.f);
        report(all->Size() >= n_touching && all->Size() == DEMSim.GetNumContacts(), "a negative threshold lists every potential pair");
        std::vector<std::pair<family_t, family_t>> fams;
        const auto pairs = DEMSim.GetContacts(fams);
        report(pairs.size() == all->Size() && fams.size() == pairs.size(), "GetContacts with family pairs has one entry per listed pair");
    }

    // ---- contact wildcards ----
    {
        DEMSim.SetFamilyContactWildcardValue(0, 1, "delta_time", 7.f);  // contacts between the two layers only
        DEMSim.DoStepDynamics();
        auto info = DEMSim.GetContactDetailedInfo(1e-6f);
        size_t n01 = 0, n01_set = 0, other_set = 0;
        for (size_t i = 0; i < info->Size(); i++) {
            const bool between = (info->GetAOwnerFamily()[i] + info->GetBOwnerFamily()[i] == 1) && info->GetContactType()[i] == "SS";
            const bool is_set = std::fabs(info->GetWildcard("delta_time")[i] - (7.f + (float)h)) < 1e-4f;
            if (between) { n01++; n01_set += is_set; }
            else other_set += is_set;
        }
        report(n01 == 16 && n01_set == 16 && other_set == 0, "SetFamilyContactWildcardValue(0, 1, delta_time) reached the 16 inter-layer contacts only");
        report(info->Size() >= 32, "the contact list survived the edit");
        bool refused = false;
        try { DEMSim.SetContactWildcardValue("no_such_word", 1.f); } catch (const std::exception&) { refused = true; }
        report(refused, "an unknown wildcard name is refused");
        refused = false;
        try { DEMSim.GetOwnerWildcardValue(0, "gran_strain"); } catch (const std::exception&) { refused = true; }
        report(refused, "owner wildcards are unknown to the built-in models");
    }

    // ---- persistent contacts ----
    {
        const size_t listed = DEMSim.GetNumContacts();
        DEMSim.MarkFamilyPersistentContact(0, 1);
        const size_t marked = DEMSim.GetNumPersistentContacts();
        report(marked >= 16 && marked < listed, "MarkFamilyPersistentContact(0, 1) marked the inter-layer pairs (" + std::to_string(marked) + ")");
        // lift the top layer out of reach: the broad phase drops the pairs, the marks keep them reported
        auto pos = bed_tracker->Positions();
        for (size_t i = 16; i < 32; i++) bed_tracker->SetPos(pos[i] + make_float3(0, 0, 0.3f), i);
        DEMSim.DoStepDynamics();
        size_t still = 0;
        for (const auto& pr : DEMSim.GetClumpContacts())
            if ((pr.first < 16) != (pr.second < 16)) still++;
        report(still == marked, "marked pairs stay listed after the broad phase dropped them (" + std::to_string(still) + ")");
        DEMSim.RemovePersistentContact();
        report(DEMSim.GetNumPersistentContacts() == 0, "RemovePersistentContact clears the marks");
        still = 0;
        for (const auto& pr : DEMSim.GetClumpContacts())
            if ((pr.first < 16) != (pr.second < 16)) still++;
        report(still == 0, "and the dropped pairs are gone from the list");
    }

    // ---- AddAcc: the next step only ----
    {
        const float3 v0 = loner_tracker->Vel();
        loner_tracker->AddAcc(make_float3(100.f, 0, 0));
        DEMSim.DoStepDynamics();
        const float3 v1 = loner_tracker->Vel();
        DEMSim.DoStepDynamics();
        const float3 v2 = loner_tracker->Vel();
        report(std::fabs((v1.x - v0.x) - 100.f * (float)h) < 1e-7f, "AddAcc adds a * h to the velocity in the next step");
        report(std::fabs(v2.x - v1.x) < 1e-9f, "and nothing in the step after");
        report(std::fabs((v2.z - v1.z) + 9.81f * (float)h) < 2e-7f, "gravity acts as before");
    }

    // ---- UpdateSimParams ----
    {
        DEMSim.SetGravitationalAcceleration(make_float3(0, 0, 0));
        DEMSim.UpdateSimParams();
        const float3 v0 = loner_tracker->Vel();
        DEMSim.DoStepDynamics();
        report(std::fabs(loner_tracker->Vel().z - v0.z) < 1e-9f, "UpdateSimParams pushes the changed gravity");
    }

    printf("%d checks failed\n", n_failed);
    return n_failed;
}
