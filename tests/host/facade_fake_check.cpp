// The deme::DEMSolver facade driven over tests/host/fake_core.cpp (a recording test double of the device-touching C ABI: no
// physics, no CPU path) -- the facade's HOST logic on a machine without a GPU: what Initialize() flattens and uploads, what
// the file writers print, the detailed contact read-out, persistent marks, wildcard edits, region inspectors, trackers.
// Prints "ok <group>" lines; a failed expectation aborts with its line number.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>

#include "../../include/dem_b200.h"

using namespace deme;

struct FakeContact {
    uint32_t a, b;
    uint8_t type;
    float wc[4], force[3], point[3];
};
extern "C" {
DemCtx* fake_last_ctx(void);
void fake_add_contact(DemCtx*, uint32_t a, uint32_t b, int type, const float wc[4], const float force[3], const float point[3]);
void fake_clear_contacts(DemCtx*);
void fake_set_reduce(DemCtx*, int kind, double v);
int fake_num_set_contacts(DemCtx*);
int fake_num_rebuilds(DemCtx*);
uint32_t fake_num_owners(DemCtx*);
uint32_t fake_num_spheres(DemCtx*);
uint32_t fake_num_triangles(DemCtx*);
uint32_t fake_num_anal(DemCtx*);
uint32_t fake_sphere_owner(DemCtx*, uint32_t s);
uint32_t fake_sphere_comp(DemCtx*, uint32_t s);
float fake_comp_radius(DemCtx*, uint32_t k);
int fake_mask(DemCtx*, unsigned i, unsigned j);
const DemPrescription* fake_prescription(DemCtx*, unsigned fam);
const DemSimParams* fake_params(DemCtx*);
double fake_option(DemCtx*, const char* name, double missing);
const float* fake_added_acc(DemCtx*);
const FakeContact* fake_set_contact(DemCtx*, uint32_t i);
uint32_t fake_num_contacts_set(DemCtx*);
float fake_material(DemCtx*, const char* what, uint32_t i, uint32_t j);
int fake_anal_type(DemCtx*, uint32_t k);
uint32_t fake_anal_owner(DemCtx*, uint32_t k);
const float* fake_anal_pos(DemCtx*, uint32_t k);
const float* fake_anal_dir(DemCtx*, uint32_t k);
float fake_anal_size(DemCtx*, uint32_t k);
uint8_t fake_owner_family(DemCtx*, uint32_t o);
}

static void expect(bool cond, const char* what, int line) {
    if (!cond) {
        fprintf(stderr, "FAILED line %d: %s\n", line, what);
        exit(1);
    }
}
#define EXPECT(c) expect((c), #c, __LINE__)
static bool throws(const std::function<void()>& f, const char* needle = nullptr) {
    try {
        f();
    } catch (const std::exception& e) {
        return !needle || std::string(e.what()).find(needle) != std::string::npos;
    }
    return false;
}
static bool close(double a, double b, double tol = 1e-6) { return std::fabs(a - b) <= tol * (1.0 + std::fabs(b)); }
static std::vector<std::vector<std::string>> read_csv(const std::string& path) {
    std::ifstream f(path);
    std::vector<std::vector<std::string>> rows;
    std::string line;
    while (std::getline(f, line)) {
        std::vector<std::string> r(1);
        for (char ch : line) {
            if (ch == ',') r.emplace_back();
            else r.back().push_back(ch);
        }
        rows.push_back(r);
    }
    return rows;
}

int main(int argc, char** argv) {
    const std::string out = argc > 1 ? argv[1] : "/tmp";
    DEMSolver sim;
    DemCtx* fake = fake_last_ctx();
    EXPECT(fake != nullptr);
    sim.SetVerbosity(QUIET);
    sim.SetOutputContent({"XYZ", "QUAT", "VEL", "ANG_VEL", "ACC", "FAMILY"});
    sim.SetContactOutputContent({"OWNER", "GEO_ID", "FORCE", "POINT", "NORMAL", "CNT_WILDCARD"});
    auto mat = sim.LoadMaterial({{"E", 1e8f}, {"nu", 0.3f}, {"CoR", 0.5f}, {"mu", 0.4f}, {"Crr", 0.0f}});
    auto ball = sim.LoadSphereType(2.f, 0.01f, mat);
    ball->AssignName("ball");
    auto pair_t = sim.LoadClumpType(3.f, make_float3(1, 2, 3), std::vector<float>{0.01f, 0.02f},
                                    std::vector<float3>{make_float3(-0.01f, 0, 0), make_float3(0.02f, 0, 0)}, mat);
    pair_t->AssignName("pair");
    sim.InstructBoxDomainDimension({-1, 1}, {-1, 1}, {0, 2});
    sim.InstructBoxDomainBoundingBC("top_open", mat);  // one external object with five planes
    // three balls (families 0, 1, 2) and two pairs (family 3), then a plane object (family 7) and a two-facet mesh (family 9)
    auto balls = sim.AddClumps(ball, std::vector<float3>{make_float3(-0.5f, 0, 0.5f), make_float3(0, 0, 0.5f), make_float3(0.5f, 0, 1.5f)});
    balls->SetFamilies(std::vector<unsigned int>{0, 1, 2});
    balls->SetVel(std::vector<float3>{make_float3(1, 0, 0), make_float3(0, 2, 0), make_float3(0, 0, -3)});
    auto pairs = sim.AddClumps(pair_t, std::vector<float3>{make_float3(-0.5f, 0.5f, 1.f), make_float3(0.5f, 0.5f, 1.f)});
    pairs->SetFamily(3);
    // the second pair is turned a quarter about z: its spheres lie along +y instead of +x
    pairs->SetOriQ(std::vector<float4>{make_float4(0, 0, 0, 1), QuatFromAxisAngle(make_float3(0, 0, 1), (float)(PI / 2))});
    pairs->SetAngVel(make_float3(0, 0, 10.f));
    auto wall = sim.AddExternalObject();
    wall->AddPlane(make_float3(0, 0, 0.1f), make_float3(0, 0, 1), mat);
    wall->SetFamily(7);
    DEMMeshConnected tri;
    tri.SetGeometry({make_float3(0, 0, 0), make_float3(1, 0, 0), make_float3(0, 1, 0), make_float3(1, 1, 0)},
                    {make_int3(0, 1, 2), make_int3(1, 3, 2)});
    tri.SetMaterial(mat);
    auto mesh = sim.AddWavefrontMeshObject(tri);
    mesh->SetFamily(9);
    mesh->SetInitPos(make_float3(0, 0, 0.2f));
    sim.SetFamilyFixed(7);
    sim.SetFamilyPrescribedLinVel(9, "0.5", "none", "0.1 * t");
    sim.DisableContactBetweenFamilies(7, 9);
    sim.DisableFamilyOutput(2);
    auto balls_tracker = sim.Track(balls);
    auto pairs_tracker = sim.Track(pairs);
    auto mesh_tracker = sim.Track(mesh);
    sim.SetInitTimeStep(1.0 / 65536.0);
    sim.SetCDUpdateFreq(-7);  // negative: adaptive off (and at least 1)
    sim.Initialize();

    {  // ---- what Initialize() flattened and uploaded ----
        // owners: 5 clumps, the plane object, the bounding box object, the mesh; spheres owner by owner
        EXPECT(fake_num_owners(fake) == 8 && sim.GetNumClumps() == 5 && sim.GetNumOwners() == 8);
        EXPECT(fake_num_spheres(fake) == 3 + 2 * 2 && fake_num_triangles(fake) == 2 && fake_num_anal(fake) == 1 + 5);
        const uint32_t want_owner[7] = {0, 1, 2, 3, 3, 4, 4};
        for (uint32_t s = 0; s < 7; s++) EXPECT(fake_sphere_owner(fake, s) == want_owner[s]);
        EXPECT(close(fake_comp_radius(fake, fake_sphere_comp(fake, 0)), 0.01) && close(fake_comp_radius(fake, fake_sphere_comp(fake, 4)), 0.02));
        EXPECT(fake_mask(fake, 7, 9) == 1 && fake_mask(fake, 9, 7) == 1 && fake_mask(fake, 0, 1) == 0);
        const DemPrescription* fixed = fake_prescription(fake, 7);
        EXPECT(fixed->used && fixed->linVelPrescribed[0] && fixed->rotVelPrescribed[2] && fixed->linVel[1] == 0.f);
        const DemPrescription* moving = fake_prescription(fake, 9);
        EXPECT(moving->used && moving->hasLinVel[0] && close(moving->linVel[0], 0.5) && !moving->hasLinVel[1] && moving->hasLinVel[2]);
        EXPECT(moving->linVel[2] == 0.f);  // 0.1 * t at t = 0
        const DemSimParams* p = fake_params(fake);
        EXPECT(p->cd_update_freq == 1 && p->force_model == DEM_HERTZIAN && p->integrator == DEM_EXTENDED_TAYLOR);
        EXPECT(p->nvXp2 + p->nvYp2 + p->nvZp2 == 64 && p->h == (float)(1.0 / 65536.0) && p->record_contact_forces == 1);
        EXPECT(fake_option(fake, "adaptive_update_freq", -1) == 0.0 && fake_option(fake, "keep_acc", -1) == 1.0);
        // read-backs go through the position code: a unit of l is far below float resolution here
        const auto pos = balls_tracker->Positions();
        EXPECT(close(pos[0].x, -0.5) && close(pos[2].z, 1.5) && close(pairs_tracker->Pos(1).y, 0.5));
        EXPECT(close(mesh_tracker->Pos().z, 0.2) && balls_tracker->GetOwnerIDs().size() == 3 && pairs_tracker->GetOwnerID(1) == 4);
        EXPECT(close(balls_tracker->Velocities()[1].y, 2) && close(pairs_tracker->AngVelLocal(0).z, 10));
        const float3 wg = pairs_tracker->AngularVelocitiesGlobal()[1];
        EXPECT(close(wg.z, 10) && close(wg.x, 0, 1e-6));
        EXPECT(balls_tracker->GetFamilies()[2] == 2 && pairs_tracker->GetFamily(1) == 3 && mesh_tracker->GetFamily() == 9);
        EXPECT(close(pairs_tracker->Mass(0), 3) && close(pairs_tracker->MOIs()[1].y, 2) && close(balls_tracker->Masses()[0], 2));
        puts("ok initialize");
    }

    {  // ---- stepping bookkeeping: the reference's step count rule, time-dependent prescriptions refreshed per step ----
        sim.DoDynamics(10.0 / 65536.0);
        EXPECT(close(sim.GetSimTime(), 10.0 / 65536.0, 1e-12));
        EXPECT(close(fake_prescription(fake, 9)->linVel[2], 0.1 * 9.0 / 65536.0, 1e-6));  // the value the LAST step started from
        sim.DoStepDynamics();
        EXPECT(close(sim.GetSimTime(), 11.0 / 65536.0, 1e-12));
        puts("ok stepping");
    }

    // contacts as the core would list them: ball 0 -- ball 1 in touch, ball 1 on the plane object (component 0), the big sphere
    // of pair 0 (sphere 4) against facet 1, and a potential pair (no force) between the two pairs
    const float w_touch[4] = {1e-4f, 2e-4f, 3e-4f, 0.25f}, w_none[4] = {0, 0, 0, 0};
    const float f01[3] = {-3, 0, 0}, p01[3] = {-0.25f, 0, 0.5f};
    const float fpl[3] = {0, 0, 5}, ppl[3] = {0, 0, 0.49f};
    const float ftr[3] = {0, 0, 7}, ptr[3] = {-0.48f, 0.5f, 0.98f};
    const float zero[3] = {0, 0, 0};
    fake_add_contact(fake, 0, 1, DEM_CNT_SPHERE_SPHERE, w_touch, f01, p01);
    fake_add_contact(fake, 1, 0, DEM_CNT_SPHERE_PLANE, w_touch, fpl, ppl);
    fake_add_contact(fake, 4, 1, DEM_CNT_SPHERE_MESH, w_touch, ftr, ptr);
    fake_add_contact(fake, 3, 5, DEM_CNT_SPHERE_SPHERE, w_none, zero, zero);

    {  // ---- detailed read-out ----
        auto info = sim.GetContactDetailedInfo(1e-6f);
        EXPECT(info->Size() == 3);  // sorted by (type, A, B): SS 0-1, SM 4-1, SA 1-0
        EXPECT(info->GetContactType()[0] == "SS" && info->GetContactType()[1] == "SM" && info->GetContactType()[2] == "SA");
        EXPECT(info->GetAOwner()[0] == 0 && info->GetBOwner()[0] == 1 && info->GetAOwnerFamily()[0] == 0 && info->GetBOwnerFamily()[0] == 1);
        EXPECT(info->GetAOwner()[1] == 3 && info->GetBOwner()[1] == 7 && info->GetBOwnerFamily()[1] == 9 && info->GetAGeo()[1] == 4);
        EXPECT(info->GetAOwner()[2] == 1 && info->GetBOwner()[2] == 5 && info->GetBOwnerFamily()[2] == 7);
        // normals: from the centre of sphere A to the contact point.  Ball 0 sits at (-0.5, 0, 0.5): the point is at +x
        EXPECT(close(info->GetNormal()[0].x, 1) && close(info->GetNormal()[0].z, 0, 1e-6));
        EXPECT(close(info->GetNormal()[2].z, -1));  // ball 1 above the plane contact point
        // sphere 4 = second component of pair 0 (offset +0.02 x, not rotated): centre (-0.48, 0.5, 1), the point is below it
        EXPECT(close(info->GetNormal()[1].z, -1) && close(info->GetNormal()[1].x, 0, 1e-5));
        EXPECT(close(info->GetForce()[2].z, 5) && close(info->GetPoint()[1].z, 0.98) && close(info->GetWildcard("delta_time")[0], 0.25));
        EXPECT(sim.GetContactDetailedInfo(-1.f)->Size() == 4 && sim.GetNumContacts() == 4);
        std::vector<std::pair<family_t, family_t>> fams;
        const auto all = sim.GetContacts(fams);
        EXPECT(all.size() == 4 && fams.size() == 4 && all[0].first == 0 && all[0].second == 1);
        EXPECT(sim.GetClumpContacts().size() == 2 && sim.GetClumpContacts(std::set<family_t>{3}).size() == 1);
        EXPECT(pairs_tracker->GetContactClumps(0).size() == 1 && pairs_tracker->GetContactClumps(0)[0] == 4);
        std::vector<float3> pts, frc, trq;
        EXPECT(balls_tracker->GetContactForces(pts, frc, 1) == 2);  // ball 1: B side of the SS contact (force negated), A side on the plane
        EXPECT(close(frc[0].x, 3) && close(frc[1].z, 5));
        EXPECT(balls_tracker->GetContactForcesAndGlobalTorque(pts, frc, trq, 0) == 1 && trq.size() == 1 && trq[0].x == 0.f);
        puts("ok contact_readout");
    }

    {  // ---- files: the reference's columns ----
        sim.WriteContactFile(out + "/fake_contacts.csv");
        auto rows = read_csv(out + "/fake_contacts.csv");
        const std::vector<std::string> header = {"contact_type", "A", "B", "geoA", "geoB", "f_x", "f_y", "f_z", "X", "Y", "Z", "n_x",
                                                 "n_y", "n_z", "delta_tan_x", "delta_tan_y", "delta_tan_z", "delta_time"};
        EXPECT(rows.size() == 4 && rows[0] == header);
        EXPECT(rows[1][0] == "SS" && rows[1][1] == "0" && rows[1][2] == "1" && rows[1][3] == "0" && rows[1][4] == "1");
        EXPECT(close(atof(rows[1][5].c_str()), -3) && close(atof(rows[1][8].c_str()), -0.25) && close(atof(rows[1][17].c_str()), 0.25));
        EXPECT(rows[2][0] == "SM" && rows[2][2] == "7" && rows[2][4] == "1" && rows[3][0] == "SA" && rows[3][2] == "5");
        sim.WriteContactFileIncludingPotentialPairs(out + "/fake_contacts_all.csv");
        EXPECT(read_csv(out + "/fake_contacts_all.csv").size() == 5);
        // the restart readers take the pairs and the history back from the file
        const auto back = DEMSolver::ReadContactPairsFromCsv(out + "/fake_contacts_all.csv");
        const auto wcs = DEMSolver::ReadContactWildcardsFromCsv(out + "/fake_contacts_all.csv");
        EXPECT(back.size() == 2 && back[0].first == 0 && back[1].second == 5 && wcs.size() == 4 && close(wcs.at("delta_tan_y")[0], 2e-4));

        sim.WriteClumpFile(out + "/fake_clumps.csv");
        rows = read_csv(out + "/fake_clumps.csv");
        const std::vector<std::string> ch = {"X", "Y", "Z", "Qw", "Qx", "Qy", "Qz", "clump_type", "v_x", "v_y", "v_z", "w_x", "w_y",
                                             "w_z", "a_x", "a_y", "a_z", "family"};
        EXPECT(rows[0] == ch && rows.size() == 1 + 4);  // family 2 is left out
        EXPECT(rows[1][7] == "ball" && rows[3][7] == "pair" && rows[2][17] == "1" && rows[4][17] == "3");
        EXPECT(close(atof(rows[2][9].c_str()), 2) && close(atof(rows[4][13].c_str()), 10) && close(atof(rows[4][3].c_str()), std::cos(PI / 4)));
        const auto xyz = DEMSolver::ReadClumpXyzFromCsv(out + "/fake_clumps.csv");
        EXPECT(xyz.at("ball").size() == 2 && xyz.at("pair").size() == 2 && close(xyz.at("pair")[1].x, 0.5));

        sim.WriteSphereFile(out + "/fake_spheres.csv");
        rows = read_csv(out + "/fake_spheres.csv");
        EXPECT(rows[0][0] == "X" && rows[0][3] == "r" && rows[0][4] == "v_x" && rows.size() == 1 + 2 + 4);
        // the turned pair: its big sphere sits at +0.02 along y of the clump centre (0.5, 0.5, 1)
        EXPECT(close(atof(rows[6][0].c_str()), 0.5, 1e-5) && close(atof(rows[6][1].c_str()), 0.52, 1e-5) && close(atof(rows[6][3].c_str()), 0.02));
        sim.WriteMeshFile(out + "/fake_mesh.vtk");
        std::ifstream vtk(out + "/fake_mesh.vtk");
        std::stringstream ss;
        ss << vtk.rdbuf();
        EXPECT(ss.str().find("POINTS 4 float") != std::string::npos && ss.str().find("CELLS 2 8") != std::string::npos);
        sim.SetMeshOutputFormat("OBJ");
        EXPECT(throws([&] { sim.WriteMeshFile(out + "/fake_mesh.obj"); }, "not implemented"));
        EXPECT(throws([&] { sim.SetOutputFormat("CHPF"); }, "ChPF") && throws([&] { sim.SetOutputFormat("xml"); }, "unknown"));
        puts("ok files");
    }

    {  // ---- contact wildcards and persistent marks ----
        sim.SetFamilyContactWildcardValue(0, 1, "delta_time", 9.f);
        EXPECT(fake_num_set_contacts(fake) == 1 && fake_num_rebuilds(fake) == 1 && fake_num_contacts_set(fake) == 4);
        int changed = 0;
        for (uint32_t i = 0; i < 4; i++) changed += fake_set_contact(fake, i)->wc[3] == 9.f;
        EXPECT(changed == 1);
        EXPECT(close(sim.GetContactDetailedInfo(1e-6f)->GetWildcard("delta_time")[0], 9));
        sim.SetFamilyContactWildcardValueEither(7, "delta_tan_x", -1.f);
        EXPECT(fake_num_set_contacts(fake) == 2 && close(sim.GetContactDetailedInfo(1e-6f)->GetWildcard("delta_tan_x")[2], -1));
        sim.SetFamilyContactWildcardValueBoth(5, "delta_tan_x", 4.f);  // nobody is in family 5: nothing to hand over
        EXPECT(fake_num_set_contacts(fake) == 2);
        EXPECT(throws([&] { sim.SetContactWildcardValue("no_such", 1.f); }, "no_such"));
        EXPECT(throws([&] { sim.SetOwnerWildcardValue(0, "gran_strain", 1.f); }, "gran_strain"));

        sim.MarkFamilyPersistentContact(3, 3);
        EXPECT(sim.GetNumPersistentContacts() == 1);
        sim.MarkFamilyPersistentContactEither(9);
        EXPECT(sim.GetNumPersistentContacts() == 2);
        fake_clear_contacts(fake);  // the broad phase proposes nothing any more
        EXPECT(sim.GetContacts().size() == 2 && sim.GetClumpContacts().size() == 1 && sim.GetContactDetailedInfo(-1.f)->Size() == 2);
        EXPECT(sim.GetContactDetailedInfo(1e-6f)->Size() == 0);
        sim.WriteContactFileIncludingPotentialPairs(out + "/fake_contacts_persistent.csv");
        EXPECT(read_csv(out + "/fake_contacts_persistent.csv").size() == 3);
        sim.RemoveFamilyPersistentContactBoth(3);
        EXPECT(sim.GetNumPersistentContacts() == 1);
        sim.RemovePersistentContact();
        EXPECT(sim.GetNumPersistentContacts() == 0 && sim.GetContacts().empty());
        puts("ok wildcards_persistence");
    }

    {  // ---- inspectors: device reductions passed through, regions and facade-side quantities evaluated here ----
        fake_set_reduce(fake, DEM_REDUCE_SPHERE_MAX_Z, 1.51);
        fake_set_reduce(fake, DEM_REDUCE_MAX_ABSV, 0.25);
        EXPECT(close(sim.CreateInspector("clump_max_z")->GetValue(), 1.51));
        // "max_absv" looks at every owner: the mesh moves at (0.5, 0, 0.1 t)
        EXPECT(close(sim.CreateInspector("max_absv")->GetValue(), 0.25));
        mesh_tracker->SetVel(make_float3(0.5f, 0, 0));
        EXPECT(close(sim.CreateInspector("max_absv")->GetValue(), 0.5));
        // regions: spheres with X < 0 are ball 0 and pair 0 (top of its big sphere at 1.02); owners with Z > 0.9: the two pairs + ball 2
        EXPECT(close(sim.CreateInspector("clump_max_z", "return X < 0;")->GetValue(), 1.02));
        EXPECT(close(sim.CreateInspector("clump_min_z", "return (X > 0.4) && (Y > 0.4);")->GetValue(), 0.98));
        EXPECT(close(sim.CreateInspector("clump_mass", "return Z > 0.9;")->GetValue(), 3 + 3 + 2));
        // ball 2 falls at 3 m/s; the turned pair spins at 10 rad/s: its small sphere (0.01 off centre) moves at 0.1 m/s
        EXPECT(close(sim.CreateInspector("clump_max_absv", "return Z > 1.2;")->GetValue(), 3));
        EXPECT(close(sim.CreateInspector("clump_max_absv", "return (X > 0.4) && (Y > 0.4) && (Z < 1.2);")->GetValue(), 0.2, 1e-5));
        const double ke_pair = 0.5 * 3.0 * 10 * 10;  // I_zz = 3
        EXPECT(close(sim.CreateInspector("clump_kinetic_energy", "return Y > 0.4;")->GetValue(), 2 * ke_pair));
        auto absv_insp = sim.CreateInspector("absv");
        float* absv = absv_insp->GetValues();  // (valid while the inspector lives)
        EXPECT(close(absv[0], 1) && close(absv[1], 2) && close(absv[2], 3) && close(absv[7], 0.5));
        puts("ok inspectors");
    }

    {  // ---- per-owner controls ----
        pairs_tracker->AddAcc(std::vector<float3>{make_float3(1, 2, 3), make_float3(4, 5, 6)});
        const float* a = fake_added_acc(fake);
        EXPECT(a[0] == 3.f && a[1] == 2.f && a[2] == 1.f && a[7] == 6.f && a[8] == 0.f);
        EXPECT(throws([&] { pairs_tracker->AddAcc(std::vector<float3>{make_float3(1, 2, 3)}); }, "tracks 2 owners"));
        balls_tracker->SetFamily(6);
        EXPECT(balls_tracker->GetFamilies()[0] == 6 && balls_tracker->GetFamilies()[2] == 6 && pairs_tracker->GetFamily(0) == 3);
        EXPECT(sim.ChangeClumpFamily(8, {-1, -0.1}, {-1, 1}, {0, 2}) == 2);  // ball 0 and pair 0 sit at X = -0.5
        EXPECT(balls_tracker->GetFamily(0) == 8 && pairs_tracker->GetFamily(0) == 8 && pairs_tracker->GetFamily(1) == 3);
        balls_tracker->SetPos(make_float3(0.25f, 0.25f, 0.75f), 1);
        EXPECT(close(balls_tracker->Pos(1).x, 0.25) && close(balls_tracker->Pos(1).z, 0.75));
        sim.SetGravitationalAcceleration(make_float3(0, 0, -1.62f));
        sim.UpdateSimParams();
        EXPECT(close(fake_params(fake)->G[2], -1.62));
        EXPECT(throws([&] { sim.ChangeFamilyWhen(0, 1, "return Z < 0;"); }, "runtime compilation"));
        EXPECT(throws([&] { sim.CorrectFamilyLinVel(0, "1", "none", "none"); }, "run-time"));
        puts("ok controls");
    }

    {  // ---- a second solver: material pair tables, analytical components, restart contacts, UpdateClumps ----
        DEMSolver sim2;
        DemCtx* f2 = fake_last_ctx();
        EXPECT(f2 != fake);
        sim2.SetVerbosity(QUIET);
        auto m1 = sim2.LoadMaterial({{"E", 1e8f}, {"nu", 0.3f}, {"CoR", 0.4f}, {"mu", 0.2f}, {"Crr", 0.0f}});
        auto m2 = sim2.LoadMaterial({{"E", 2e8f}, {"nu", 0.2f}, {"CoR", 0.8f}, {"mu", 0.6f}, {"Crr", 0.1f}});
        sim2.SetMaterialPropertyPair("CoR", m1, m2, 0.9f);
        auto b1 = sim2.LoadSphereType(1.f, 0.05f, m1);
        auto b2 = sim2.LoadSphereType(4.f, 0.1f, m2);
        sim2.InstructBoxDomainDimension(4, 4, 4);
        sim2.InstructBoxDomainBoundingBC("all", m2);
        auto first = sim2.AddClumps(b1, std::vector<float3>{make_float3(0, 0, 0), make_float3(0.1f, 0, 0)});
        first->SetExistingContacts({{0, 1}});
        first->SetExistingContactWildcards({{"delta_tan_x", {1e-5f}}, {"delta_tan_y", {2e-5f}}, {"delta_tan_z", {3e-5f}}, {"delta_time", {0.5f}}});
        auto second = sim2.AddClumps(b2, make_float3(1, 1, 1));
        second->SetFamily(4);
        auto can = sim2.AddExternalObject();
        can->AddCylinder(make_float3(0, 0, 0), make_float3(0, 0, 2), 1.5f, m1, ENTITY_NORMAL_INWARD);
        can->AddPlane(make_float3(0, 0, -1), make_float3(0, 0, 3), m1);
        can->SetFamily(20);
        sim2.SetFamilyFixed(20);
        auto can_tracker = sim2.Track(can);
        auto first_tracker = sim2.Track(first);
        sim2.SetInitTimeStep(1e-5);
        sim2.SetCDUpdateFreq(15);
        sim2.UseAdaptiveUpdateFreq(false);
        sim2.Initialize();
        // materials: pairwise properties default to the mean unless set (APIPrivate.cpp:1944 of the reference)
        EXPECT(close(fake_material(f2, "E", 1, 0), 2e8) && close(fake_material(f2, "nu", 0, 0), 0.3));
        EXPECT(close(fake_material(f2, "CoR", 0, 0), 0.4) && close(fake_material(f2, "CoR", 0, 1), 0.9) && close(fake_material(f2, "CoR", 1, 0), 0.9));
        EXPECT(close(fake_material(f2, "mu", 0, 1), 0.4) && close(fake_material(f2, "Crr", 1, 0), 0.05) && close(fake_material(f2, "Crr", 1, 1), 0.1));
        // analytical components: the user's cylinder and plane (normals normalised), then the six planes of the box
        EXPECT(fake_num_anal(f2) == 2 + 6 && fake_anal_type(f2, 0) == DEM_ANAL_CYL_INF && fake_anal_type(f2, 1) == DEM_ANAL_PLANE);
        EXPECT(close(fake_anal_dir(f2, 0)[2], 1) && close(fake_anal_size(f2, 0), 1.5) && close(fake_anal_dir(f2, 1)[2], 1) && close(fake_anal_pos(f2, 1)[2], -1));
        EXPECT(fake_anal_owner(f2, 0) == 3 && fake_anal_owner(f2, 1) == 3 && fake_anal_owner(f2, 2) == 4 && fake_anal_owner(f2, 7) == 4);
        int up = 0, down = 0;
        for (uint32_t k = 2; k < 8; k++) {
            EXPECT(fake_anal_type(f2, k) == DEM_ANAL_PLANE);
            up += fake_anal_dir(f2, k)[2] > 0.5f && close(fake_anal_pos(f2, k)[2], -2);
            down += fake_anal_dir(f2, k)[2] < -0.5f && close(fake_anal_pos(f2, k)[2], 2);
        }
        EXPECT(up == 1 && down == 1);
        EXPECT(fake_owner_family(f2, 3) == 20 && fake_owner_family(f2, 4) == RESERVED_FAMILY_NUM && can_tracker->GetOwnerID() == 3);
        EXPECT(fake_params(f2)->cd_update_freq == 15 && fake_option(f2, "adaptive_update_freq", -1) == 0.0);
        // the restart contacts went in as the history source of the first rebuild
        EXPECT(fake_num_set_contacts(f2) == 1 && fake_num_contacts_set(f2) == 1);
        EXPECT(fake_set_contact(f2, 0)->a == 0 && fake_set_contact(f2, 0)->b == 1 && fake_set_contact(f2, 0)->type == DEM_CNT_SPHERE_SPHERE);
        EXPECT(close(fake_set_contact(f2, 0)->wc[1], 2e-5) && close(fake_set_contact(f2, 0)->wc[3], 0.5));

        // ... the run goes on: things move, contacts exist; then more clumps are added on the fly
        const float wc[4] = {1e-4f, 0, 0, 0.75f}, fz[3] = {0, 0, 1}, pt[3] = {0.05f, 0, 0};
        fake_add_contact(f2, 0, 1, DEM_CNT_SPHERE_SPHERE, wc, fz, pt);
        fake_add_contact(f2, 2, 0, DEM_CNT_SPHERE_CYL, wc, fz, pt);
        sim2.DoDynamics(5e-5);
        first_tracker->SetPos(make_float3(0.3f, 0.2f, 0.1f), 1);
        first_tracker->SetVel(make_float3(-1, -2, -3), 1);
        first_tracker->SetFamily(11, 1);
        const double t_before = sim2.GetSimTime();
        auto late = sim2.AddClumps(b2, std::vector<float3>{make_float3(-1, -1, 1), make_float3(-1, 1, 1)});
        late->SetFamily(5);
        late->SetVel(make_float3(0, 0, -1));
        auto late_tracker = sim2.Track(late);
        sim2.UpdateClumps();
        // the newcomers are numbered right behind the existing clumps; everything that was there keeps its state
        EXPECT(sim2.GetNumClumps() == 5 && sim2.GetNumOwners() == 7 && fake_num_owners(f2) == 7 && fake_num_spheres(f2) == 5);
        EXPECT(close(first_tracker->Pos(1).x, 0.3) && close(first_tracker->Vel(1).z, -3) && first_tracker->GetFamily(1) == 11);
        EXPECT(late_tracker->GetOwnerID(0) == 3 && close(late_tracker->Pos(1).y, 1) && close(late_tracker->Vel(0).z, -1) && late_tracker->GetFamily(1) == 5);
        EXPECT(can_tracker->GetOwnerID() == 5 && fake_owner_family(f2, 5) == 20 && fake_anal_owner(f2, 0) == 5 && fake_anal_owner(f2, 2) == 6);
        EXPECT(close(sim2.GetSimTime(), t_before, 1e-12));
        // and the contact list of before went back in, pair by pair, with its history
        EXPECT(fake_num_set_contacts(f2) == 2 && fake_num_contacts_set(f2) == 2);
        bool ss = false, cyl = false;
        for (uint32_t i = 0; i < 2; i++) {
            const FakeContact* k = fake_set_contact(f2, i);
            ss = ss || (k->type == DEM_CNT_SPHERE_SPHERE && k->a == 0 && k->b == 1 && close(k->wc[3], 0.75));
            cyl = cyl || (k->type == DEM_CNT_SPHERE_CYL && k->a == 2 && k->b == 0 && close(k->wc[0], 1e-4));
        }
        EXPECT(ss && cyl);
        puts("ok second_solver");
    }
    return 0;
}
