// fake_core.cpp -- a TEST DOUBLE of the device-touching part of the C ABI (include/dem_b200.h).  Not product code and not a
// CPU path: it computes NO physics.  It records what the deme::DEMSolver facade uploads and answers read-backs from that
// record plus tables a test injects (contacts, reductions), so that the facade's host-side logic -- flattening at
// Initialize(), file writers, contact read-outs, persistent marks, region inspectors, trackers -- can run under ASan on a
// machine without a GPU.  Linked IN FRONT of libdemcore.so by tests/test_facade_fake_core_cpu.py only; the host-only
// entry points (dem_host_*) still come from the real library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dem_b200.h"

struct FakeContact {
    uint32_t a, b;
    uint8_t type;
    float wc[4], force[3], point[3];
};

struct DemCtx {
    DemSimParams sp{};
    bool params_set = false, initialized = false;
    std::string err;
    // what was uploaded
    std::vector<float> radii, relX, relY, relZ, mass, moiX, moiY, moiZ;
    uint32_t nMat = 0;
    std::vector<float> E, nu, CoR, mu, Crr;
    std::vector<uint32_t> analOwner;
    std::vector<uint8_t> analType;
    std::vector<float> analNormalSign, analPos, analDir, analSize1;
    std::vector<uint8_t> masks;
    std::vector<float> extra;
    std::vector<DemPrescription> presc;
    int n_family_uploads = 0;
    std::vector<uint64_t> voxel;
    std::vector<uint16_t> lx, ly, lz, inertia;
    std::vector<float> quat, vel, omg, acc, angacc;  // quat w,x,y,z
    std::vector<uint8_t> family;
    std::vector<uint32_t> sphOwner;
    std::vector<uint16_t> sphComp, sphMat;
    std::vector<uint32_t> triOwner;
    // what the test injects / the facade sets
    std::vector<FakeContact> contacts, contacts_set;
    int n_set_contacts = 0, n_rebuilds = 0;
    double reduce_value[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<std::pair<std::string, double>> options;
    std::vector<float> added_acc;  // first, n, then values: last dem_add_owner_acc
    uint64_t n_steps = 0;
    double sim_time = 0;
    uint32_t nClumps() const {
        uint32_t n = 0;
        for (uint32_t o : sphOwner) n = std::max(n, o + 1);
        return n;
    }
};

static DemCtx* g_last = nullptr;

static int fail(DemCtx* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}

extern "C" {

// ---- controls for the test (not part of the ABI) ----
DemCtx* fake_last_ctx(void) { return g_last; }
void fake_add_contact(DemCtx* c, uint32_t a, uint32_t b, int type, const float wc[4], const float force[3], const float point[3]) {
    FakeContact k;
    k.a = a; k.b = b; k.type = (uint8_t)type;
    memcpy(k.wc, wc, sizeof(k.wc)); memcpy(k.force, force, sizeof(k.force)); memcpy(k.point, point, sizeof(k.point));
    c->contacts.push_back(k);
}
void fake_clear_contacts(DemCtx* c) { c->contacts.clear(); }
void fake_set_reduce(DemCtx* c, int kind, double v) { c->reduce_value[kind] = v; }
int fake_num_set_contacts(DemCtx* c) { return c->n_set_contacts; }
int fake_num_rebuilds(DemCtx* c) { return c->n_rebuilds; }
int fake_num_family_uploads(DemCtx* c) { return c->n_family_uploads; }
uint32_t fake_num_owners(DemCtx* c) { return (uint32_t)c->voxel.size(); }
uint32_t fake_num_spheres(DemCtx* c) { return (uint32_t)c->sphOwner.size(); }
uint32_t fake_num_triangles(DemCtx* c) { return (uint32_t)c->triOwner.size(); }
uint32_t fake_num_anal(DemCtx* c) { return (uint32_t)c->analOwner.size(); }
uint32_t fake_sphere_owner(DemCtx* c, uint32_t s) { return c->sphOwner.at(s); }
uint32_t fake_sphere_comp(DemCtx* c, uint32_t s) { return c->sphComp.at(s); }
float fake_comp_radius(DemCtx* c, uint32_t k) { return c->radii.at(k); }
int fake_mask(DemCtx* c, unsigned i, unsigned j) {
    if (i > j) std::swap(i, j);
    return c->masks.at((1 + j) * j / 2 + i);
}
const DemPrescription* fake_prescription(DemCtx* c, unsigned fam) { return &c->presc.at(fam); }
const DemSimParams* fake_params(DemCtx* c) { return &c->sp; }
double fake_option(DemCtx* c, const char* name, double missing) {
    double v = missing;
    for (const auto& kv : c->options)
        if (kv.first == name) v = kv.second;
    return v;
}
const float* fake_added_acc(DemCtx* c) { return c->added_acc.data(); }
const FakeContact* fake_set_contact(DemCtx* c, uint32_t i) { return &c->contacts_set.at(i); }
float fake_material(DemCtx* c, const char* what, uint32_t i, uint32_t j) {
    const std::string w = what;
    if (w == "E") return c->E.at(i);
    if (w == "nu") return c->nu.at(i);
    const std::vector<float>& t = (w == "CoR") ? c->CoR : (w == "mu" ? c->mu : c->Crr);
    return t.at((size_t)i * c->nMat + j);
}
int fake_anal_type(DemCtx* c, uint32_t k) { return c->analType.at(k); }
uint32_t fake_anal_owner(DemCtx* c, uint32_t k) { return c->analOwner.at(k); }
const float* fake_anal_pos(DemCtx* c, uint32_t k) { return &c->analPos.at(3 * k); }
const float* fake_anal_dir(DemCtx* c, uint32_t k) { return &c->analDir.at(3 * k); }
float fake_anal_size(DemCtx* c, uint32_t k) { return c->analSize1.at(k); }
uint8_t fake_owner_family(DemCtx* c, uint32_t o) { return c->family.at(o); }
uint32_t fake_num_contacts_set(DemCtx* c) { return (uint32_t)c->contacts_set.size(); }

// ---- the ABI ----
int dem_device_count(void) { return 1; }
int dem_ctx_create(DemCtx** out, int) {
    if (!out) return DEM_ERR_INVALID;
    *out = g_last = new DemCtx();
    return DEM_OK;
}
int dem_ctx_create_group(DemCtx** out, const int*, int) { return dem_ctx_create(out, 0); }
int dem_ctx_destroy(DemCtx* c) {
    if (g_last == c) g_last = nullptr;
    delete c;
    return DEM_OK;
}
const char* dem_last_error(const DemCtx* c) { return c ? c->err.c_str() : "no context"; }
int dem_set_params(DemCtx* c, const DemSimParams* p) {
    if (!c || !p) return DEM_ERR_INVALID;
    if (p->cd_update_freq < 1) return fail(c, DEM_ERR_INVALID, "cd_update_freq must be >= 1");
    c->sp = *p;
    c->params_set = true;
    return DEM_OK;
}
int dem_upload_templates(DemCtx* c, uint32_t nComp, const float* radii, const float* relX, const float* relY, const float* relZ,
                         uint32_t nMassProps, const float* mass, const float* moiX, const float* moiY, const float* moiZ) {
    c->radii.assign(radii, radii + nComp); c->relX.assign(relX, relX + nComp);
    c->relY.assign(relY, relY + nComp); c->relZ.assign(relZ, relZ + nComp);
    c->mass.assign(mass, mass + nMassProps); c->moiX.assign(moiX, moiX + nMassProps);
    c->moiY.assign(moiY, moiY + nMassProps); c->moiZ.assign(moiZ, moiZ + nMassProps);
    return DEM_OK;
}
int dem_upload_materials(DemCtx* c, uint32_t nMat, const float* E, const float* nu, const float* CoR, const float* mu,
                         const float* Crr) {
    c->nMat = nMat;
    c->E.assign(E, E + nMat); c->nu.assign(nu, nu + nMat);
    c->CoR.assign(CoR, CoR + (size_t)nMat * nMat); c->mu.assign(mu, mu + (size_t)nMat * nMat);
    c->Crr.assign(Crr, Crr + (size_t)nMat * nMat);
    return DEM_OK;
}
int dem_upload_analytical(DemCtx* c, uint32_t nAnal, const uint32_t* objOwner, const uint8_t* objType, const uint16_t*,
                          const float* objNormal, const float* relPosX, const float* relPosY, const float* relPosZ,
                          const float* rotX, const float* rotY, const float* rotZ, const float* size1, const float*, const float*,
                          const float*) {
    c->analOwner.assign(objOwner, objOwner + nAnal);
    c->analType.assign(objType, objType + nAnal);
    c->analNormalSign.assign(objNormal, objNormal + nAnal);
    c->analSize1.assign(size1, size1 + nAnal);
    c->analPos.clear(); c->analDir.clear();
    for (uint32_t i = 0; i < nAnal; i++) {
        c->analPos.insert(c->analPos.end(), {relPosX[i], relPosY[i], relPosZ[i]});
        c->analDir.insert(c->analDir.end(), {rotX[i], rotY[i], rotZ[i]});
    }
    return DEM_OK;
}
int dem_upload_families(DemCtx* c, const uint8_t* masks, const float* extra, const DemPrescription* presc) {
    c->masks.assign(masks, masks + DEM_NUM_FAMILY_MASKS);
    c->extra.assign(extra, extra + DEM_NUM_FAMILIES);
    c->presc.assign(presc, presc + DEM_NUM_FAMILIES);
    c->n_family_uploads++;
    return DEM_OK;
}
int dem_upload_owners(DemCtx* c, uint32_t n, const uint64_t* voxelID, const uint16_t* locX, const uint16_t* locY,
                      const uint16_t* locZ, const float* qw, const float* qx, const float* qy, const float* qz, const float* vX,
                      const float* vY, const float* vZ, const float* oX, const float* oY, const float* oZ, const uint8_t* fam,
                      const uint16_t* inertia) {
    c->voxel.assign(voxelID, voxelID + n); c->lx.assign(locX, locX + n); c->ly.assign(locY, locY + n); c->lz.assign(locZ, locZ + n);
    c->quat.resize(4 * (size_t)n); c->vel.resize(3 * (size_t)n); c->omg.resize(3 * (size_t)n);
    c->acc.assign(3 * (size_t)n, 0.f); c->angacc.assign(3 * (size_t)n, 0.f);
    for (uint32_t i = 0; i < n; i++) {
        c->quat[4 * i] = qw[i]; c->quat[4 * i + 1] = qx[i]; c->quat[4 * i + 2] = qy[i]; c->quat[4 * i + 3] = qz[i];
        c->vel[3 * i] = vX[i]; c->vel[3 * i + 1] = vY[i]; c->vel[3 * i + 2] = vZ[i];
        c->omg[3 * i] = oX[i]; c->omg[3 * i + 1] = oY[i]; c->omg[3 * i + 2] = oZ[i];
    }
    c->family.assign(fam, fam + n);
    c->inertia.assign(inertia, inertia + n);
    return DEM_OK;
}
int dem_upload_spheres(DemCtx* c, uint32_t n, const uint32_t* owner, const uint16_t* comp, const uint16_t* mat) {
    c->sphOwner.assign(owner, owner + n); c->sphComp.assign(comp, comp + n); c->sphMat.assign(mat, mat + n);
    return DEM_OK;
}
int dem_upload_triangles(DemCtx* c, uint32_t n, const uint32_t* ownerMesh, const float*, const float*, const float*, const uint16_t*) {
    if (n) c->triOwner.assign(ownerMesh, ownerMesh + n); else c->triOwner.clear();
    return DEM_OK;
}
int dem_update_triangle_nodes(DemCtx*, uint32_t, uint32_t, const float*, const float*, const float*) { return DEM_OK; }
int dem_initialize(DemCtx* c, uint64_t) {
    if (!c->params_set) return fail(c, DEM_ERR_INVALID, "dem_set_params must precede dem_initialize");
    c->initialized = true;
    return DEM_OK;
}
int dem_set_option(DemCtx* c, const char* name, double value) {
    c->options.emplace_back(name, value);
    return DEM_OK;
}
int dem_set_contacts(DemCtx* c, uint64_t n, const uint32_t* idA, const uint32_t* idB, const uint8_t* type, const float* wc4) {
    c->contacts_set.clear();
    for (uint64_t i = 0; i < n; i++) {
        FakeContact k{};
        k.a = idA[i]; k.b = idB[i]; k.type = type[i];
        if (wc4) memcpy(k.wc, wc4 + 4 * i, sizeof(k.wc));
        c->contacts_set.push_back(k);
    }
    c->n_set_contacts++;
    return DEM_OK;
}
int dem_rebuild_contacts(DemCtx* c) {
    // what a rebuild does with a list handed over by dem_set_contacts: the pairs stay, with their history
    if (!c->contacts_set.empty()) {
        for (auto& k : c->contacts)
            for (const auto& s : c->contacts_set)
                if (s.a == k.a && s.b == k.b && s.type == k.type) memcpy(k.wc, s.wc, sizeof(k.wc));
    }
    c->n_rebuilds++;
    return DEM_OK;
}
static void advance(DemCtx* c, uint64_t n) {
    c->n_steps += n;
    c->sim_time += (double)n * (double)c->sp.h;
}
int dem_do_dynamics(DemCtx* c, double t) {
    uint64_t n = 0;
    for (double cycle = 0.0; cycle < t; cycle += (double)c->sp.h) n++;
    advance(c, n);
    return DEM_OK;
}
int dem_step(DemCtx* c, uint64_t n) { advance(c, n); return DEM_OK; }
int dem_step_async(DemCtx* c, uint64_t n) { advance(c, n); return DEM_OK; }
int dem_sync(DemCtx*) { return DEM_OK; }
int dem_set_sim_time(DemCtx* c, double t) { c->sim_time = t; return DEM_OK; }
int dem_update_step_size(DemCtx* c, float h) { c->sp.h = h; return DEM_OK; }

int dem_download_owner_state(DemCtx* c, uint32_t first, uint32_t n, uint64_t* voxelID, uint16_t* locX, uint16_t* locY,
                             uint16_t* locZ, float* q, float* vel, float* omg, float* acc, float* angacc, uint8_t* family) {
    if ((uint64_t)first + n > c->voxel.size()) return fail(c, DEM_ERR_INVALID, "owner range out of bounds");
    for (uint32_t i = 0; i < n; i++) {
        const size_t o = first + i;
        if (voxelID) voxelID[i] = c->voxel[o];
        if (locX) locX[i] = c->lx[o];
        if (locY) locY[i] = c->ly[o];
        if (locZ) locZ[i] = c->lz[o];
        if (q) memcpy(q + 4 * i, &c->quat[4 * o], 16);
        if (vel) memcpy(vel + 3 * i, &c->vel[3 * o], 12);
        if (omg) memcpy(omg + 3 * i, &c->omg[3 * o], 12);
        if (acc) memcpy(acc + 3 * i, &c->acc[3 * o], 12);
        if (angacc) memcpy(angacc + 3 * i, &c->angacc[3 * o], 12);
        if (family) family[i] = c->family[o];
    }
    return DEM_OK;
}
int dem_download_positions(DemCtx* c, uint32_t first, uint32_t n, float* x32, double* x64) {
    if ((uint64_t)first + n > c->voxel.size()) return fail(c, DEM_ERR_INVALID, "owner range out of bounds");
    const DemSimParams& p = c->sp;
    for (uint32_t i = 0; i < n; i++) {
        const size_t o = first + i;
        const uint64_t vx = c->voxel[o] & ((1ull << p.nvXp2) - 1ull), vy = (c->voxel[o] >> p.nvXp2) & ((1ull << p.nvYp2) - 1ull),
                       vz = c->voxel[o] >> (p.nvXp2 + p.nvYp2);
        const double X[3] = {(double)vx * p.voxelSize + (double)c->lx[o] * p.l, (double)vy * p.voxelSize + (double)c->ly[o] * p.l,
                             (double)vz * p.voxelSize + (double)c->lz[o] * p.l};
        for (int k = 0; k < 3; k++) {
            if (x64) x64[3 * i + k] = X[k] + (double)p.LBF[k];
            if (x32) x32[3 * i + k] = (float)(X[k] + (double)p.LBF[k]);
        }
    }
    return DEM_OK;
}
int dem_upload_owner_state(DemCtx* c, uint32_t first, uint32_t n, const float* pos, const float* q, const float* vel,
                           const float* omg, const uint8_t* family) {
    if ((uint64_t)first + n > c->voxel.size()) return fail(c, DEM_ERR_INVALID, "owner range out of bounds");
    if (pos) {
        std::vector<uint64_t> v(n);
        std::vector<uint16_t> a(n), b(n), d(n);
        dem_host_encode_positions(&c->sp, pos, n, v.data(), a.data(), b.data(), d.data());  // (the real host routine)
        for (uint32_t i = 0; i < n; i++) { c->voxel[first + i] = v[i]; c->lx[first + i] = a[i]; c->ly[first + i] = b[i]; c->lz[first + i] = d[i]; }
    }
    for (uint32_t i = 0; i < n; i++) {
        const size_t o = first + i;
        if (q) memcpy(&c->quat[4 * o], q + 4 * i, 16);
        if (vel) memcpy(&c->vel[3 * o], vel + 3 * i, 12);
        if (omg) memcpy(&c->omg[3 * o], omg + 3 * i, 12);
        if (family) c->family[o] = family[i];
    }
    return DEM_OK;
}
int dem_add_owner_acc(DemCtx* c, uint32_t first, uint32_t n, const float* acc, const float* angacc) {
    c->added_acc.assign({(float)first, (float)n});
    for (uint32_t i = 0; i < 3 * n; i++) c->added_acc.push_back(acc ? acc[i] : 0.f);
    for (uint32_t i = 0; i < 3 * n; i++) c->added_acc.push_back(angacc ? angacc[i] : 0.f);
    return DEM_OK;
}
int dem_set_family_material(DemCtx*, uint32_t, uint32_t, int) { return DEM_OK; }
int dem_download_contact_records(DemCtx* c, uint64_t capacity, uint64_t* n, uint32_t* idA, uint32_t* idB, uint8_t* type,
                                 float* wc4, float* force, float* point) {
    // the real library hands the rows out sorted by (type, idA, idB)
    std::vector<FakeContact> rows = c->contacts;
    std::sort(rows.begin(), rows.end(), [](const FakeContact& x, const FakeContact& y) {
        if (x.type != y.type) return x.type < y.type;
        if (x.a != y.a) return x.a < y.a;
        return x.b < y.b;
    });
    *n = rows.size();
    if (!idA && !idB && !type && !wc4 && !force && !point) return DEM_OK;
    if (capacity < rows.size()) return fail(c, DEM_ERR_CAPACITY, "need room for more contacts");
    for (size_t i = 0; i < rows.size(); i++) {
        if (idA) idA[i] = rows[i].a;
        if (idB) idB[i] = rows[i].b;
        if (type) type[i] = rows[i].type;
        if (wc4) memcpy(wc4 + 4 * i, rows[i].wc, 16);
        if (force) memcpy(force + 3 * i, rows[i].force, 12);
        if (point) memcpy(point + 3 * i, rows[i].point, 12);
    }
    return DEM_OK;
}
int dem_download_contacts(DemCtx* c, uint64_t capacity, uint64_t* n, uint32_t* idA, uint32_t* idB, uint8_t* type, float* wc4,
                          float* force) {
    return dem_download_contact_records(c, capacity, n, idA, idB, type, wc4, force, nullptr);
}
int dem_get_stats(DemCtx* c, DemStats* out) {
    memset(out, 0, sizeof(*out));
    out->n_steps = c->n_steps;
    out->sim_time = c->sim_time;
    out->cd_update_freq = c->sp.cd_update_freq;
    for (const auto& k : c->contacts) {
        if (k.type == DEM_CNT_SPHERE_SPHERE) out->n_contacts_ss++;
        else if (k.type == DEM_CNT_SPHERE_MESH) out->n_contacts_st++;
        else out->n_contacts_sa++;
    }
    return DEM_OK;
}
int dem_reduce(DemCtx* c, int kind, double* out) {
    if (kind < 0 || kind > 7) return fail(c, DEM_ERR_INVALID, "unknown reduction");
    *out = c->reduce_value[kind];
    return DEM_OK;
}

}  // extern "C"
