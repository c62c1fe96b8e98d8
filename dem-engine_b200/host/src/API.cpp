// API.cpp -- implementation of the deme::DEMSolver facade over the C ABI of include/dem_b200.h.
// Flattening follows DEMSolver::Initialize of the reference (src/DEM/APIPublic.cpp:2161-2213,
// APIPrivate.cpp:119-1120, dT.cpp:638-1024): owners ordered clumps, then external objects; the world bounding box is
// one extra external object appended at Initialize(); materials pairwise-averaged unless set explicitly.
#include <DEM/API.h>
#include <set>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "../../../include/dem_b200.h"

namespace deme {

namespace {
std::string g_data_path = "";

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error(msg); }

// "none" means unspecified. Anything else is an expression of the simulation time t (DEM/utils/Expression.hpp): a
// constant is folded here; a time-dependent one is kept in `expr` and re-evaluated before every step.
bool parse_prescription(const std::string& s, double t_now, float& out, std::shared_ptr<TimeExpression>& expr) {
    std::string t;
    for (char c : s)
        if (!isspace((unsigned char)c)) t.push_back(c);
    expr.reset();
    if (t.empty() || t == "none") return false;
    auto e = std::make_shared<TimeExpression>(s);
    out = (float)e->Eval(t_now);
    if (!e->IsConstant()) expr = e;
    return true;
}

std::string upper(std::string s) {
    for (auto& c : s) c = (char)toupper((unsigned char)c);
    return s;
}
}  // namespace

std::filesystem::path GET_DATA_PATH() {
    if (!g_data_path.empty()) return std::filesystem::path(g_data_path);
    if (const char* e = std::getenv("DEME_DATA_PATH")) return std::filesystem::path(e);
    return std::filesystem::path("data");
}
std::filesystem::path GetDEMEDataFile(const std::string& relative) { return GET_DATA_PATH() / relative; }
void SetDEMEDataPath(const std::string& path) { g_data_path = path; }

// ---------------------------------------------------------------------------------------------------------------
int DEMClumpTemplate::ReadComponentFromFile(const std::string filename, const std::string x_id, const std::string y_id,
                                            const std::string z_id, const std::string r_id) {
    std::ifstream f(filename);
    if (!f) fail("Clump template file " + filename + " cannot be opened.");
    std::string line;
    std::getline(f, line);
    std::vector<std::string> cols;
    {
        std::stringstream ss(line);
        std::string c;
        while (std::getline(ss, c, ',')) {
            c.erase(std::remove_if(c.begin(), c.end(), [](unsigned char ch) { return isspace(ch); }), c.end());
            cols.push_back(c);
        }
    }
    // A column the file does not have reads as 0 for every row (the reference opens the file with ignore_missing_column,
    // Structs.h:629-648: data/clumps/ViperWheelSimple.csv has no radii and DEMdemo_WheelSlopeSlip sets them afterwards), and
    // the rows are APPENDED to what the template already holds.
    const size_t none = (size_t)-1;
    auto col = [&](const std::string& id) {
        auto it = std::find(cols.begin(), cols.end(), id);
        return it == cols.end() ? none : (size_t)(it - cols.begin());
    };
    const size_t ix = col(x_id), iy = col(y_id), iz = col(z_id), ir = col(r_id);
    unsigned int count = 0;
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::stringstream ss(line);
        std::string c;
        std::vector<float> v;
        while (std::getline(ss, c, ',')) v.push_back((float)std::atof(c.c_str()));
        auto at = [&](size_t i) { return (i == none || i >= v.size()) ? 0.f : v[i]; };
        relPos.push_back(make_float3(at(ix), at(iy), at(iz)));
        radii.push_back(at(ir));
        count++;
    }
    nComp += count;
    return 0;
}

void DEMClumpTemplate::Scale(float s) {
    if (!(s > 0)) fail("Scale: s must be positive");
    for (auto& pos : relPos) pos *= s;
    for (auto& rad : radii) rad *= s;
    const double ps = (double)std::abs(s);
    mass *= ps * ps * ps;
    MOI *= ps * ps * ps * ps * ps;
    volume *= ps * ps * ps;
}

DEMClumpBatch::DEMClumpBatch(size_t num) : nClumps(num) {
    types.resize(num);
    families.resize(num, 0);
    vel.resize(num, make_float3(0, 0, 0));
    angVel.resize(num, make_float3(0, 0, 0));
    xyz.resize(num);
    oriQ.resize(num, make_float4(0, 0, 0, 1));
    obj_type = OWNER_TYPE::CLUMP;
}
void DEMClumpBatch::assertLength(size_t len, const std::string& name) const {
    if (len != nClumps) {
        std::stringstream ss;
        ss << name << " input argument must have length " << nClumps << " (not " << len
           << "), same as the number of clumps you originally added via AddClumps." << std::endl;
        throw std::runtime_error(ss.str());
    }
}
void DEMClumpBatch::SetTypes(const std::vector<std::shared_ptr<DEMClumpTemplate>>& input) { assertLength(input.size(), "SetTypes"); types = input; }
void DEMClumpBatch::SetPos(const std::vector<float3>& input) { assertLength(input.size(), "SetPos"); xyz = input; }
void DEMClumpBatch::SetVel(const std::vector<float3>& input) { assertLength(input.size(), "SetVel"); vel = input; }
void DEMClumpBatch::SetAngVel(const std::vector<float3>& input) { assertLength(input.size(), "SetAngVel"); angVel = input; }
void DEMClumpBatch::SetOriQ(const std::vector<float4>& input) { assertLength(input.size(), "SetOriQ"); oriQ = input; }
void DEMClumpBatch::SetFamilies(const std::vector<unsigned int>& input) {
    assertLength(input.size(), "SetFamilies");
    for (unsigned int f : input)
        if (f > 255) fail("A clump is instructed to have a family number larger than the max allowance 255");
    families = input;
    family_isSpecified = true;
}

void DEMClumpBatch::AddExistingContactWildcard(const std::string& name, const std::vector<float>& vals) {
    if (vals.size() != contact_pairs.size())
        fail("AddExistingContactWildcard needs one value per existing contact (" + std::to_string(contact_pairs.size()) +
             " set via SetExistingContacts), but " + std::to_string(vals.size()) + " were given for " + name + ".");
    contact_wildcards[name] = vals;
}
void DEMClumpBatch::SetOwnerWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards) {
    for (const auto& kv : wildcards) assertLength(kv.second.size(), "SetOwnerWildcards");
    owner_wildcards = wildcards;
}
void DEMClumpBatch::AddOwnerWildcard(const std::string& name, const std::vector<float>& vals) {
    assertLength(vals.size(), "AddOwnerWildcard");
    owner_wildcards[name] = vals;
}
void DEMClumpBatch::SetGeometryWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards) {
    for (const auto& kv : wildcards)
        if (kv.second.size() != nSpheres)
            fail("SetGeometryWildcards needs one value per sphere of the batch (" + std::to_string(nSpheres) + "), not " +
                 std::to_string(kv.second.size()) + ".");
    geo_wildcards = wildcards;
}
void DEMClumpBatch::AddGeometryWildcard(const std::string& name, const std::vector<float>& vals) {
    if (vals.size() != nSpheres)
        fail("AddGeometryWildcard needs one value per sphere of the batch (" + std::to_string(nSpheres) + "), not " +
             std::to_string(vals.size()) + ".");
    geo_wildcards[name] = vals;
}

void DEMExternObj::SetFamily(const unsigned int code) {
    if (code > 255) fail("An external object is instructed to have a family number larger than the max allowance 255");
    family_code = code;
}
void DEMExternObj::AddPlane(const float3 pos, const float3 normal, const std::shared_ptr<DEMMaterial>& material) {
    comps.push_back({0, pos, normalize(normal), 0.f, 0.f, material});
}
void DEMExternObj::AddZCylinder(const float3 pos, const float rad, const std::shared_ptr<DEMMaterial>& material,
                                const objNormal_t normal) {
    comps.push_back({2, pos, make_float3(0, 0, 1), rad, normal ? 1.f : 0.f, material});
}
void DEMExternObj::AddCylinder(const float3 pos, const float3 axis, const float rad,
                               const std::shared_ptr<DEMMaterial>& material, const objNormal_t normal) {
    comps.push_back({2, pos, normalize(axis), rad, normal ? 1.f : 0.f, material});
}

// ---------------------------------------------------------------------------------------------------------------
// DEMSolver(nGPUs) (src/DEM/API.h:53; APIPublic.cpp:23-74): the reference gives each of its two worker threads a GPU.  Here
// nGPUs devices (as many as the box has, at most 8) form ONE group context: a scene large enough to be worth it is sharded
// into x-slabs over them, anything else runs on the first device.  DEME_B200_GPUS overrides the count, DEME_B200_DEVICE
// picks the (first) device.
static DemCtx* create_group_ctx(std::vector<int> ids) {
    const int have = dem_device_count();
    if (have <= 0) fail("DEMSolver: no usable CUDA device (this core has no CPU fallback)");
    std::vector<int> use;
    for (int d : ids)
        if (d >= 0 && d < have && std::find(use.begin(), use.end(), d) == use.end() && use.size() < 8) use.push_back(d);
    if (use.empty()) use.push_back(0);
    DemCtx* c = nullptr;
    const int rc = dem_ctx_create_group(&c, use.data(), (int)use.size());
    if (rc != DEM_OK) fail("DEMSolver: dem_ctx_create_group returned " + std::to_string(rc) + " (this core has no CPU fallback)");
    return c;
}
DEMSolver::DEMSolver(unsigned int nGPUs) {
    int first = 0, n = (int)nGPUs;
    if (const char* e = std::getenv("DEME_B200_DEVICE")) first = std::atoi(e);
    if (const char* e = std::getenv("DEME_B200_GPUS")) n = std::atoi(e);
    std::vector<int> ids;
    for (int k = 0; k < std::max(n, 1); k++) ids.push_back(first + k);
    ctx = create_group_ctx(ids);
    m_family_masks.assign(DEM_NUM_FAMILY_MASKS, 0);
    if (const char* e = std::getenv("DEME_B200_ADAPTIVE_FREQ")) m_adaptive_update_freq = std::atoi(e) != 0;
}
DEMSolver::DEMSolver(const std::vector<int>& gpu_ids) {
    ctx = create_group_ctx(gpu_ids);
    m_family_masks.assign(DEM_NUM_FAMILY_MASKS, 0);
    if (const char* e = std::getenv("DEME_B200_ADAPTIVE_FREQ")) m_adaptive_update_freq = std::atoi(e) != 0;
}
DEMSolver::~DEMSolver() {
    if (ctx) dem_ctx_destroy(ctx);
}

void DEMSolver::check(int rc, const char* what) const {
    if (rc != DEM_OK) fail(std::string(what) + " failed: " + dem_last_error(ctx));
}
void DEMSolver::assertInit(const char* what) const {
    if (!sys_initialized) fail(std::string(what) + " can only be called after Initialize().");
}

void DEMSolver::SetVerbosity(const std::string& verbose) {
    const std::string u = upper(verbose);
    if (u == "QUIET") verbosity = QUIET; else if (u == "ERROR") verbosity = ERR; else if (u == "WARNING") verbosity = WARNING;
    else if (u == "INFO") verbosity = INFO; else if (u == "STEP_ANOMALY") verbosity = STEP_ANOMALY;
    else if (u == "STEP_METRIC") verbosity = STEP_METRIC; else if (u == "DEBUG") verbosity = DEBUG;
    else if (u == "STEP_DEBUG") verbosity = STEP_DEBUG;
    else fail("Instruction " + verbose + " is unknown in SetVerbosity call.");
}
namespace {
OUTPUT_FORMAT parse_output_format(const std::string& format, const char* who) {
    const std::string u = upper(format);
    if (u == "CSV") return OUTPUT_FORMAT::CSV;
    if (u == "BINARY") return OUTPUT_FORMAT::BINARY;
    if (u == "CHPF") fail("ChPF is not enabled when the code was compiled.");
    fail("Instruction " + format + " is unknown in " + who + " call.");
}
}  // namespace
void DEMSolver::SetOutputFormat(OUTPUT_FORMAT format) {
    if (format == OUTPUT_FORMAT::CHPF) fail("ChPF is not enabled when the code was compiled.");
    m_out_format = format;
}
void DEMSolver::SetOutputFormat(const std::string& format) { m_out_format = parse_output_format(format, "SetOutputFormat"); }
void DEMSolver::SetContactOutputFormat(OUTPUT_FORMAT format) {
    if (format == OUTPUT_FORMAT::CHPF) fail("ChPF is not enabled when the code was compiled.");
    m_cnt_out_format = format;
}
void DEMSolver::SetContactOutputFormat(const std::string& format) {
    m_cnt_out_format = parse_output_format(format, "SetContactOutputFormat");
}
void DEMSolver::SetMeshOutputFormat(const std::string& format) {
    const std::string u = upper(format);
    if (u == "VTK") m_mesh_out_format = MESH_FORMAT::VTK;
    else if (u == "OBJ") m_mesh_out_format = MESH_FORMAT::OBJ;
    else fail("Instruction " + format + " is unknown in SetMeshOutputFormat call.");
}
void DEMSolver::warnIfBinary(OUTPUT_FORMAT f, const char* what) const {
    if (f == OUTPUT_FORMAT::BINARY && verbosity >= WARNING)
        std::cerr << "WARNING! Binary " << what << " output is not implemented yet, using CSV..." << std::endl;
}
void DEMSolver::SetOutputContent(const std::vector<std::string>& content) {
    unsigned int c = XYZ;
    for (const auto& a : content) {
        const std::string u = upper(a);
        if (u == "XYZ") c |= XYZ; else if (u == "QUAT") c |= QUAT; else if (u == "ABSV") c |= ABSV; else if (u == "VEL") c |= VEL;
        else if (u == "ANG_VEL") c |= ANG_VEL; else if (u == "ABS_ACC") c |= ABS_ACC; else if (u == "ACC") c |= ACC;
        else if (u == "ANG_ACC") c |= ANG_ACC; else if (u == "FAMILY") c |= FAMILY; else if (u == "MAT") c |= MAT;
        else fail("Instruction " + a + " is unknown in SetOutputContent call.");
    }
    m_out_content = c;
}
void DEMSolver::SetContactOutputContent(const std::vector<std::string>& content) {
    unsigned int c = CNT_TYPE;
    for (const auto& a : content) {
        const std::string u = upper(a);
        if (u == "CNT_TYPE") c |= CNT_TYPE; else if (u == "FORCE") c |= FORCE; else if (u == "POINT" || u == "CNT_POINT") c |= CNT_POINT;
        else if (u == "COMPONENT") c |= COMPONENT; else if (u == "NORMAL") c |= NORMAL; else if (u == "TORQUE") c |= TORQUE;
        else if (u == "CNT_WILDCARD") c |= CNT_WILDCARD; else if (u == "OWNER") c |= OWNER; else if (u == "GEO_ID") c |= GEO_ID;
        else fail("Instruction " + a + " is unknown in SetContactOutputContent call.");
    }
    m_cnt_out_content = c;
}

namespace {
int exact_axis(const std::string& dir_exact) {
    const std::string u = upper(dir_exact);
    if (u == "X") return 0;
    if (u == "Y") return 1;
    if (u == "Z") return 2;
    if (u == "NONE") return -1;
    fail("Unknown '" + dir_exact + "' parameter in InstructBoxDomainDimension call.");
}
}  // namespace
void DEMSolver::InstructBoxDomainDimension(float x, float y, float z, const std::string& dir_exact) {
    InstructBoxDomainDimension({-x / 2.f, x / 2.f}, {-y / 2.f, y / 2.f}, {-z / 2.f, z / 2.f}, dir_exact);
    if (exact_axis(dir_exact) < 0) {  // (the centred form goes through the C ABI's own arithmetic)
        float umin[3], umax[3], tmin[3], tmax[3];
        dem_host_box_domain(x, y, z, umin, umax, tmin, tmax);
        m_user_box_min = make_float3(umin[0], umin[1], umin[2]);
        m_user_box_max = make_float3(umax[0], umax[1], umax[2]);
        m_target_box_min = make_float3(tmin[0], tmin[1], tmin[2]);
        m_target_box_max = make_float3(tmax[0], tmax[1], tmax[2]);
    }
}
void DEMSolver::InstructBoxDomainDimension(const std::pair<float, float>& x, const std::pair<float, float>& y,
                                           const std::pair<float, float>& z, const std::string& dir_exact) {
    m_box_dir_exact = exact_axis(dir_exact);
    // APIPublic.cpp:874-905: enlarge by 20 % about the user's box, except along the axis whose length is to be exact
    m_user_box_min = make_float3(std::min(x.first, x.second), std::min(y.first, y.second), std::min(z.first, z.second));
    m_user_box_max = make_float3(std::max(x.first, x.second), std::max(y.first, y.second), std::max(z.first, z.second));
    const float3 sz = m_user_box_max - m_user_box_min;
    float3 enl = make_float3(sz.x * 0.2f / 2.f, sz.y * 0.2f / 2.f, sz.z * 0.2f / 2.f);
    if (m_box_dir_exact == 0) enl.x = 0.f;
    if (m_box_dir_exact == 1) enl.y = 0.f;
    if (m_box_dir_exact == 2) enl.z = 0.f;
    m_target_box_min = m_user_box_min - enl;
    m_target_box_max = m_user_box_max + enl;
}
void DEMSolver::InstructBoxDomainBoundingBC(const std::string& inst, const std::shared_ptr<DEMMaterial>& mat) {
    if (inst != "none" && inst != "all" && inst != "top_open" && inst != "only_bottom" && inst != "only_sides")
        fail("Domain bounding BC instruction " + inst + " is unknown.");
    m_user_add_bounding_box = inst;
    m_bounding_box_material = mat;
}

void DEMSolver::SetCDUpdateFreq(int freq) {
    // SetCDUpdateFreq(0) is the reference's lock-step mode: rebuild before every step; a negative value also switches the
    // adaptive frequency off (API.h:107-113 of the reference)
    m_cd_update_freq = std::max(1, freq);
    if (freq < 0) UseAdaptiveUpdateFreq(false);
}
void DEMSolver::SetExpandSafetyType(const std::string& insp_type) {
    if (insp_type != "auto") fail("Unknown string input \"" + insp_type + "\" for SetExpandSafetyType.");
}
float DEMSolver::GetUpdateFreq() const {
    if (!sys_initialized) return (float)m_cd_update_freq;
    DemStats st;
    dem_get_stats(ctx, &st);
    return (float)st.cd_update_freq;
}
void DEMSolver::UseAdaptiveUpdateFreq(bool flag) {
    m_adaptive_update_freq = flag;
    if (sys_initialized) check(dem_set_option(ctx, "adaptive_update_freq", flag ? 1.0 : 0.0), "UseAdaptiveUpdateFreq");
}
void DEMSolver::SetCDMaxUpdateFreq(unsigned int max_freq) {
    m_max_update_freq = std::max(1u, max_freq);
    if (sys_initialized) check(dem_set_option(ctx, "update_freq_max", (double)m_max_update_freq), "SetCDMaxUpdateFreq");
}
void DEMSolver::SetIntegrator(const std::string& intg) {
    const std::string u = upper(intg);
    if (u == "FORWARD_EULER") m_integrator = TIME_INTEGRATOR::FORWARD_EULER;
    else if (u == "CENTERED_DIFFERENCE") m_integrator = TIME_INTEGRATOR::CENTERED_DIFFERENCE;
    else if (u == "EXTENDED_TAYLOR") m_integrator = TIME_INTEGRATOR::EXTENDED_TAYLOR;
    else fail("Integration type " + intg + " is unknown. Please select another via SetIntegrator.");
}
void DEMSolver::SetExpandFactor(float beta, bool fix) {
    if (fix) m_expand_factor = beta;
}
void DEMSolver::UpdateStepSize(double ts) {
    if (ts > 0) m_ts_size = ts;
    if (sys_initialized) check(dem_update_step_size(ctx, (float)m_ts_size), "UpdateStepSize");
}

std::shared_ptr<DEMForceModel> DEMSolver::UseFrictionalHertzianModel() {
    m_force_model = FORCE_MODEL::HERTZIAN;
    m_force_model_obj = std::make_shared<DEMForceModel>(m_force_model);
    // what the model reads from the materials (AuxClasses.cpp:757-764 of the reference)
    m_force_model_obj->SetMustHaveMatProp({"E", "nu", "CoR", "mu", "Crr"});
    m_force_model_obj->SetMustPairwiseMatProp({"CoR", "mu", "Crr"});
    return m_force_model_obj;
}
std::shared_ptr<DEMForceModel> DEMSolver::UseFrictionlessHertzianModel() {
    m_force_model = FORCE_MODEL::HERTZIAN_FRICTIONLESS;
    m_force_model_obj = std::make_shared<DEMForceModel>(m_force_model);
    m_force_model_obj->SetMustHaveMatProp({"E", "nu", "CoR"});
    m_force_model_obj->SetMustPairwiseMatProp({"CoR"});
    return m_force_model_obj;
}
namespace {
const char* const kNoRuntimeCompilation =
    ": custom force-model source needs run-time compilation, which this ahead-of-time compiled core does not have. Use "
    "UseFrictionalHertzianModel() or UseFrictionlessHertzianModel().";
std::vector<std::string> history_names(FORCE_MODEL m) {
    if (m == FORCE_MODEL::HERTZIAN) return {"delta_tan_x", "delta_tan_y", "delta_tan_z", "delta_time"};
    return {};
}
}  // namespace
void DEMForceModel::SetPerContactWildcards(const std::set<std::string>& wildcards) {
    const std::vector<std::string> own = history_names(type);
    if (wildcards != std::set<std::string>(own.begin(), own.end()))
        fail("SetPerContactWildcards: the built-in force models keep their own history words (delta_tan_x, delta_tan_y, "
             "delta_tan_z, delta_time for the frictional model, none for the frictionless one); other sets belong to "
             "custom models" + std::string(kNoRuntimeCompilation));
}
void DEMForceModel::SetPerOwnerWildcards(const std::set<std::string>& wildcards) {
    if (!wildcards.empty()) fail("SetPerOwnerWildcards" + std::string(kNoRuntimeCompilation));
}
void DEMForceModel::SetPerGeometryWildcards(const std::set<std::string>& wildcards) {
    if (!wildcards.empty()) fail("SetPerGeometryWildcards" + std::string(kNoRuntimeCompilation));
}
void DEMForceModel::SetForceModelType(FORCE_MODEL model_type) {
    if (model_type == FORCE_MODEL::CUSTOM) fail("SetForceModelType(CUSTOM)" + std::string(kNoRuntimeCompilation));
    type = model_type;
}
void DEMForceModel::DefineCustomModel(const std::string&) { fail("DefineCustomModel" + std::string(kNoRuntimeCompilation)); }
int DEMForceModel::ReadCustomModelFile(const std::filesystem::path&) { fail("ReadCustomModelFile" + std::string(kNoRuntimeCompilation)); }
void DEMForceModel::DefineCustomModelPrerequisites(const std::string&) {
    fail("DefineCustomModelPrerequisites" + std::string(kNoRuntimeCompilation));
}
int DEMForceModel::ReadCustomModelPrerequisitesFile(const std::filesystem::path&) {
    fail("ReadCustomModelPrerequisitesFile" + std::string(kNoRuntimeCompilation));
}
std::shared_ptr<DEMForceModel> DEMSolver::DefineContactForceModel(const std::string&) {
    fail("DefineContactForceModel: custom force-model source needs runtime compilation, which this ahead-of-time compiled "
         "core does not have. Use UseFrictionalHertzianModel() or UseFrictionlessHertzianModel().");
}
std::shared_ptr<DEMForceModel> DEMSolver::ReadContactForceModel(const std::string&) {
    fail("ReadContactForceModel: custom force-model source needs runtime compilation, which this ahead-of-time compiled "
         "core does not have. Use UseFrictionalHertzianModel() or UseFrictionlessHertzianModel().");
}

std::shared_ptr<DEMMaterial> DEMSolver::LoadMaterial(const std::unordered_map<std::string, float>& mat_prop) {
    auto m = std::make_shared<DEMMaterial>(mat_prop);
    m->load_order = (unsigned int)m_loaded_materials.size();
    m_loaded_materials.push_back(m);
    return m;
}
void DEMSolver::SetMaterialPropertyPair(const std::string& name, const std::shared_ptr<DEMMaterial>& mat1,
                                        const std::shared_ptr<DEMMaterial>& mat2, float val) {
    m_pairwise_matprop[name][{mat1->load_order, mat2->load_order}] = val;
}

std::shared_ptr<DEMClumpTemplate> DEMSolver::LoadClumpType(DEMClumpTemplate& clump) {
    if (clump.mass <= 0 || length(clump.MOI) <= 0)
        std::cerr << "WARNING! A type of clump is instructed to have near-zero mass or moment of inertia." << std::endl;
    if (clump.radii.size() != clump.relPos.size() || clump.radii.size() != clump.materials.size())
        fail("Arrays defining a clump topology type must all have the same length.");
    auto p = std::make_shared<DEMClumpTemplate>(clump);
    p->nComp = (unsigned int)clump.radii.size();
    p->mark = (unsigned int)m_templates.size();
    if (p->m_name == "NULL") {
        char name[16];
        snprintf(name, sizeof(name), "%04u", p->mark);
        p->m_name = name;
    }
    m_templates.push_back(p);
    return p;
}
std::shared_ptr<DEMClumpTemplate> DEMSolver::LoadClumpType(float mass, float3 moi, const std::vector<float>& sp_radii,
                                                           const std::vector<float3>& sp_locations_xyz,
                                                           const std::vector<std::shared_ptr<DEMMaterial>>& sp_materials) {
    DEMClumpTemplate c;
    c.mass = mass; c.MOI = moi; c.radii = sp_radii; c.relPos = sp_locations_xyz; c.materials = sp_materials;
    c.nComp = (unsigned int)sp_radii.size();
    return LoadClumpType(c);
}
std::shared_ptr<DEMClumpTemplate> DEMSolver::LoadClumpType(float mass, float3 moi, const std::vector<float>& sp_radii,
                                                           const std::vector<float3>& sp_locations_xyz,
                                                           const std::shared_ptr<DEMMaterial>& sp_material) {
    return LoadClumpType(mass, moi, sp_radii, sp_locations_xyz,
                         std::vector<std::shared_ptr<DEMMaterial>>(sp_radii.size(), sp_material));
}
std::shared_ptr<DEMClumpTemplate> DEMSolver::LoadClumpType(float mass, float3 moi, const std::string filename,
                                                           const std::vector<std::shared_ptr<DEMMaterial>>& sp_materials) {
    DEMClumpTemplate c;
    c.mass = mass; c.MOI = moi;
    c.ReadComponentFromFile(filename);
    c.materials = sp_materials;
    return LoadClumpType(c);
}
std::shared_ptr<DEMClumpTemplate> DEMSolver::LoadClumpType(float mass, float3 moi, const std::string filename,
                                                           const std::shared_ptr<DEMMaterial>& sp_material) {
    DEMClumpTemplate c;
    c.mass = mass; c.MOI = moi;
    c.ReadComponentFromFile(filename);
    c.materials.assign(c.nComp, sp_material);
    return LoadClumpType(c);
}
std::shared_ptr<DEMClumpTemplate> DEMSolver::LoadSphereType(float mass, float radius, const std::shared_ptr<DEMMaterial>& material) {
    const float I = (float)(2.0 / 5.0 * mass * radius * radius);  // APIPublic.cpp LoadSphereType
    return LoadClumpType(mass, make_float3(I, I, I), std::vector<float>(1, radius), std::vector<float3>(1, make_float3(0, 0, 0)),
                         std::vector<std::shared_ptr<DEMMaterial>>(1, material));
}

std::shared_ptr<DEMClumpBatch> DEMSolver::AddClumps(DEMClumpBatch& input_batch) {
    auto b = std::make_shared<DEMClumpBatch>(input_batch);
    b->load_order = (unsigned int)m_cached_input_clump_batches.size();
    size_t nsp = 0;
    for (const auto& t : b->types) {
        if (!t) fail("AddClumps: a clump has no template assigned.");
        nsp += t->nComp;
    }
    b->nSpheres = nsp;
    m_cached_input_clump_batches.push_back(b);
    return b;
}
std::shared_ptr<DEMClumpBatch> DEMSolver::AddClumps(const std::vector<std::shared_ptr<DEMClumpTemplate>>& input_types,
                                                    const std::vector<float3>& input_xyz) {
    if (input_types.size() != input_xyz.size())
        fail("Arrays in the call AddClumps must all have the same length.");
    DEMClumpBatch b(input_xyz.size());
    b.SetTypes(input_types);
    b.SetPos(input_xyz);
    return AddClumps(b);
}
std::shared_ptr<DEMExternObj> DEMSolver::AddExternalObject() {
    auto o = std::make_shared<DEMExternObj>();
    o->load_order = (unsigned int)m_cached_extern_objs.size();
    m_cached_extern_objs.push_back(o);
    return o;
}
std::shared_ptr<DEMExternObj> DEMSolver::AddBCPlane(const float3 pos, const float3 normal, const std::shared_ptr<DEMMaterial>& material) {
    auto o = AddExternalObject();
    o->SetFamily(RESERVED_FAMILY_NUM);
    o->AddPlane(pos, normal, material);
    return o;
}

// ---- DEMMeshConnected (src/DEM/BdrsAndObjs.h:222-520, src/DEM/BdrsAndObjs.cpp LoadWavefrontMesh) ----
bool DEMMeshConnected::LoadWavefrontMesh(const std::string& input_file, bool load_normals, bool load_uv) {
    std::ifstream in(input_file);
    if (!in.is_open()) {
        std::cerr << "Cannot open mesh file " << input_file << std::endl;
        return false;
    }
    Clear();
    filename = input_file;
    std::string line;
    // one face corner "v", "v/vt", "v//vn" or "v/vt/vn"; negative indices count from the end
    auto corner = [&](const std::string& tok, int& v, int& vt, int& vn) {
        v = vt = vn = 0;
        size_t a = tok.find('/');
        v = std::stoi(tok.substr(0, a));
        if (a != std::string::npos) {
            size_t b = tok.find('/', a + 1);
            const std::string st = tok.substr(a + 1, b == std::string::npos ? std::string::npos : b - a - 1);
            if (!st.empty()) vt = std::stoi(st);
            if (b != std::string::npos && b + 1 < tok.size()) vn = std::stoi(tok.substr(b + 1));
        }
        if (v < 0) v = (int)m_vertices.size() + v + 1;
        if (vt < 0) vt = (int)m_UV.size() + vt + 1;
        if (vn < 0) vn = (int)m_normals.size() + vn + 1;
    };
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v") {
            float x, y, z;
            ss >> x >> y >> z;
            m_vertices.push_back(make_float3(x, y, z));
        } else if (tag == "vn") {
            float x, y, z;
            ss >> x >> y >> z;
            if (load_normals) m_normals.push_back(make_float3(x, y, z));
        } else if (tag == "vt") {
            float u = 0, v = 0;
            ss >> u >> v;
            if (load_uv) m_UV.push_back(make_float3(u, v, 0));
        } else if (tag == "f") {
            std::vector<int> vi, ti, ni;
            std::string tok;
            while (ss >> tok) {
                int v, vt, vn;
                corner(tok, v, vt, vn);
                vi.push_back(v - 1); ti.push_back(vt - 1); ni.push_back(vn - 1);
            }
            for (size_t k = 1; k + 1 < vi.size(); k++) {  // fan triangulation of polygons
                m_face_v_indices.push_back(make_int3(vi[0], vi[k], vi[k + 1]));
                if (load_normals && ni[0] >= 0) m_face_n_indices.push_back(make_int3(ni[0], ni[k], ni[k + 1]));
                if (load_uv && ti[0] >= 0) m_face_uv_indices.push_back(make_int3(ti[0], ti[k], ti[k + 1]));
            }
        }
    }
    for (const auto& fc : m_face_v_indices)
        if (fc.x < 0 || fc.y < 0 || fc.z < 0 || (size_t)fc.x >= m_vertices.size() || (size_t)fc.y >= m_vertices.size() ||
            (size_t)fc.z >= m_vertices.size())
            fail("Mesh file " + input_file + " has a facet that refers to a vertex it does not define.");
    nTri = m_face_v_indices.size();
    return true;
}
void DEMMeshConnected::SetGeometry(const std::vector<float3>& vertices, const std::vector<int3>& faces) {
    Clear();
    m_vertices = vertices;
    m_face_v_indices = faces;
    for (const auto& fc : faces)
        if (fc.x < 0 || fc.y < 0 || fc.z < 0 || (size_t)fc.x >= vertices.size() || (size_t)fc.y >= vertices.size() ||
            (size_t)fc.z >= vertices.size())
            fail("SetGeometry: a facet refers to a vertex that does not exist.");
    nTri = faces.size();
}
std::vector<std::vector<float>> DEMMeshConnected::GetCoordsVerticesAsVectorOfVectors() {
    std::vector<std::vector<float>> out;
    out.reserve(m_vertices.size());
    for (const float3& v : m_vertices) out.push_back({v.x, v.y, v.z});
    return out;
}
std::vector<std::vector<int>> DEMMeshConnected::GetIndicesVertexesAsVectorOfVectors() {
    std::vector<std::vector<int>> out;
    out.reserve(m_face_v_indices.size());
    for (const int3& f : m_face_v_indices) out.push_back({f.x, f.y, f.z});
    return out;
}
void DEMMeshConnected::WriteWavefront(const std::string& filename, std::vector<DEMMeshConnected>& meshes) {
    std::ofstream mf(filename);
    if (!mf) fail("WriteWavefront: " + filename + " cannot be opened for writing.");
    // Wavefront indices count from 1 and run through the whole file: first all nodes, then all normals, then the faces
    std::vector<size_t> v_base, n_base;
    size_t nv = 1, nn = 1;
    for (const auto& m : meshes) {
        v_base.push_back(nv);
        nv += m.m_vertices.size();
        for (const float3& v : m.m_vertices) mf << "v " << v.x << " " << v.y << " " << v.z << "\n";
    }
    for (const auto& m : meshes) {
        n_base.push_back(nn);
        nn += m.m_normals.size();
        for (const float3& v : m.m_normals) mf << "vn " << v.x << " " << v.y << " " << v.z << "\n";
    }
    for (size_t k = 0; k < meshes.size(); k++) {
        const auto& m = meshes[k];
        const bool with_normals = !m.m_normals.empty() && m.m_face_n_indices.size() == m.m_face_v_indices.size();
        for (size_t j = 0; j < m.m_face_v_indices.size(); j++) {
            const int3 f = m.m_face_v_indices[j];
            if (with_normals) {
                const int3 g = m.m_face_n_indices[j];
                mf << "f " << f.x + v_base[k] << "//" << g.x + n_base[k] << " " << f.y + v_base[k] << "//" << g.y + n_base[k]
                   << " " << f.z + v_base[k] << "//" << g.z + n_base[k] << "\n";
            } else {
                mf << "f " << f.x + v_base[k] << " " << f.y + v_base[k] << " " << f.z + v_base[k] << "\n";
            }
        }
    }
}
DEMMeshConnected DEMMeshConnected::Merge(std::vector<DEMMeshConnected>& meshes) {
    DEMMeshConnected out;
    bool all_have_materials = !meshes.empty();
    for (const auto& m : meshes) {
        const int vb = (int)out.m_vertices.size(), nb = (int)out.m_normals.size(), ub = (int)out.m_UV.size();
        out.m_vertices.insert(out.m_vertices.end(), m.m_vertices.begin(), m.m_vertices.end());
        out.m_normals.insert(out.m_normals.end(), m.m_normals.begin(), m.m_normals.end());
        out.m_UV.insert(out.m_UV.end(), m.m_UV.begin(), m.m_UV.end());
        for (const int3& f : m.m_face_v_indices) out.m_face_v_indices.push_back(make_int3(f.x + vb, f.y + vb, f.z + vb));
        for (const int3& f : m.m_face_n_indices) out.m_face_n_indices.push_back(make_int3(f.x + nb, f.y + nb, f.z + nb));
        for (const int3& f : m.m_face_uv_indices) out.m_face_uv_indices.push_back(make_int3(f.x + ub, f.y + ub, f.z + ub));
        all_have_materials = all_have_materials && m.isMaterialSet && m.materials.size() == m.nTri;
    }
    out.nTri = out.m_face_v_indices.size();
    if (out.m_face_n_indices.size() != out.nTri) { out.m_normals.clear(); out.m_face_n_indices.clear(); }  // (only some had normals)
    if (out.m_face_uv_indices.size() != out.nTri) { out.m_UV.clear(); out.m_face_uv_indices.clear(); }
    if (all_have_materials) {
        for (const auto& m : meshes) out.materials.insert(out.materials.end(), m.materials.begin(), m.materials.end());
        out.isMaterialSet = true;
    }
    return out;
}
void DEMMeshConnected::SetGeometryWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards) {
    for (const auto& kv : wildcards)
        if (kv.second.size() != nTri)
            fail("Input gemometry wildcard arrays in a SetGeometryWildcards call must all have the same size as the number of "
                 "triangles in this mesh.\nHere, the input array has length " + std::to_string(kv.second.size()) +
                 " but this mesh has " + std::to_string(nTri) + " triangles.");
    geo_wildcards = wildcards;
}
void DEMMeshConnected::AddGeometryWildcard(const std::string& name, const std::vector<float>& vals) {
    if (vals.size() != nTri)
        fail("Input gemometry wildcard array in a AddGeometryWildcard call must have the same size as the number of triangles "
             "in this mesh.\nHere, the input array has length " + std::to_string(vals.size()) + " but this mesh has " +
             std::to_string(nTri) + " triangles.");
    geo_wildcards[name] = vals;
}
void DEMMeshConnected::Clear() {
    m_vertices.clear(); m_normals.clear(); m_UV.clear(); m_colors.clear();
    m_face_v_indices.clear(); m_face_n_indices.clear(); m_face_uv_indices.clear(); m_face_col_indices.clear();
    geo_wildcards.clear();
    materials.clear(); isMaterialSet = false;
    nTri = 0;
}
void DEMMeshConnected::SetMaterial(const std::vector<std::shared_ptr<DEMMaterial>>& input) {
    if (input.size() != nTri) {
        std::stringstream ss;
        ss << "SetMaterial input argument must have length " << nTri << " (not " << input.size()
           << "), same as the number of triangle facets in the mesh." << std::endl;
        throw std::runtime_error(ss.str());
    }
    materials = input;
    isMaterialSet = true;
}
// applyFrameTransformGlobalToLocal / LocalToGlobal, src/DEM/HostSideHelpers.hpp
void DEMMeshConnected::InformCentroidPrincipal(float3 center, float4 prin_Q) {
    const float4 inv = make_float4(-prin_Q.x, -prin_Q.y, -prin_Q.z, prin_Q.w);
    for (auto& node : m_vertices) node = Rotate(node - center, inv);
}
void DEMMeshConnected::Move(float3 vec, float4 rot_Q) {
    for (auto& node : m_vertices) node = Rotate(node, rot_Q) + vec;
}
void DEMMeshConnected::Mirror(float3 plane_point, float3 plane_normal) {
    plane_normal = normalize(plane_normal);
    for (auto& node : m_vertices) node += 2.f * dot(plane_point - node, plane_normal) * plane_normal;
    for (auto& nrm : m_normals) nrm -= 2.f * dot(nrm, plane_normal) * plane_normal;
    for (auto* faces : {&m_face_v_indices, &m_face_n_indices, &m_face_uv_indices})
        for (auto& fc : *faces) std::swap(fc.y, fc.z);  // keep the winding outward after the reflection
}
void DEMMeshConnected::Scale(float s) {
    if (!(s > 0.f)) fail("Scale: the scaling factor must be positive.");
    for (auto& node : m_vertices) node *= s;
    const double d = (double)s;
    mass = (float)(mass * d * d * d);
    MOI *= d * d * d * d * d;
}
void DEMMeshConnected::Scale(float3 s) {
    if (!(s.x > 0.f && s.y > 0.f && s.z > 0.f)) fail("Scale: the scaling factors must be positive.");
    for (auto& node : m_vertices) node = node * s;
    const double prod = (double)s.x * (double)s.y * (double)s.z;
    mass = (float)(mass * prod);
    MOI.x = (float)(MOI.x * prod * s.x * s.x);
    MOI.y = (float)(MOI.y * prod * s.y * s.y);
    MOI.z = (float)(MOI.z * prod * s.z * s.z);
}

std::shared_ptr<DEMMeshConnected> DEMSolver::AddWavefrontMeshObject(DEMMeshConnected& mesh) {
    if (mesh.GetNumTriangles() == 0 && verbosity >= WARNING)
        std::cerr << "WARNING! It seems that a mesh contains 0 triangle facet at the time it is loaded." << std::endl;
    if (sys_initialized) fail("AddWavefrontMeshObject: adding meshes to an initialised system is not supported; add them before Initialize().");
    auto m = std::make_shared<DEMMeshConnected>(mesh);
    m->load_order = (unsigned int)m_cached_meshes.size();
    m_cached_meshes.push_back(m);
    return m;
}
std::shared_ptr<DEMMeshConnected> DEMSolver::AddWavefrontMeshObject(const std::string& filename, const std::shared_ptr<DEMMaterial>& mat,
                                                                    bool load_normals, bool load_uv) {
    DEMMeshConnected mesh;
    if (!mesh.LoadWavefrontMesh(filename, load_normals, load_uv)) fail("Failed to load in mesh file " + filename + ".");
    mesh.SetMaterial(mat);
    return AddWavefrontMeshObject(mesh);
}
std::shared_ptr<DEMMeshConnected> DEMSolver::AddWavefrontMeshObject(const std::string& filename, bool load_normals, bool load_uv) {
    DEMMeshConnected mesh;
    if (!mesh.LoadWavefrontMesh(filename, load_normals, load_uv)) fail("Failed to load in mesh file " + filename + ".");
    return AddWavefrontMeshObject(mesh);
}

void DEMSolver::WriteMeshFile(const std::filesystem::path& outfilename) const {
    assertInit("WriteMeshFile");
    if (m_mesh_out_format != MESH_FORMAT::VTK)
        fail("Mesh output file format is unknown or not implemented. Please re-set it via SetMeshOutputFormat.");
    std::ofstream f(outfilename);
    size_t nV = 0, nF = 0;
    for (const auto& m : m_cached_meshes) { nV += m->m_vertices.size(); nF += m->nTri; }
    f << "# vtk DataFile Version 2.0\nmeshes\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS " << nV << " float\n";
    for (const auto& m : m_cached_meshes) {
        const float3 p = GetOwnerPosition(m->owner)[0];
        const float4 q = GetOwnerOriQ(m->owner)[0];
        for (const auto& v : m->m_vertices) {
            const float3 w = Rotate(v, q) + p;
            f << w.x << " " << w.y << " " << w.z << "\n";
        }
    }
    f << "CELLS " << nF << " " << 4 * nF << "\n";
    size_t base = 0;
    for (const auto& m : m_cached_meshes) {
        for (const auto& fc : m->m_face_v_indices) f << "3 " << base + fc.x << " " << base + fc.y << " " << base + fc.z << "\n";
        base += m->m_vertices.size();
    }
    f << "CELL_TYPES " << nF << "\n";
    for (size_t i = 0; i < nF; i++) f << "5\n";
}

std::shared_ptr<DEMInspector> DEMSolver::CreateInspector(const std::string& quantity) {
    return std::make_shared<DEMInspector>(this, quantity);
}

void DEMSolver::DisableContactBetweenFamilies(unsigned int ID1, unsigned int ID2) {
    if (ID1 > 255 || ID2 > 255) fail("Family numbers must not exceed 255.");
    const unsigned i = std::min(ID1, ID2), j = std::max(ID1, ID2);
    m_family_masks[(1 + j) * j / 2 + i] = 1;
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::EnableContactBetweenFamilies(unsigned int ID1, unsigned int ID2) {
    if (ID1 > 255 || ID2 > 255) fail("Family numbers must not exceed 255.");
    const unsigned i = std::min(ID1, ID2), j = std::max(ID1, ID2);
    m_family_masks[(1 + j) * j / 2 + i] = 0;
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::SetFamilyFixed(unsigned int ID) {
    if (ID > 255) fail("Family numbers must not exceed 255.");
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    for (int k = 0; k < 3; k++) {
        p.linVelP[k] = p.rotVelP[k] = p.linPosP[k] = true;
        p.hasLinVel[k] = p.hasRotVel[k] = true;
        p.linVel[k] = p.rotVel[k] = 0.f;
    }
    p.rotPosP = true;
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::SetFamilyClumpMaterial(unsigned int N, const std::shared_ptr<DEMMaterial>& mat) {
    assertInit("SetFamilyClumpMaterial");
    check(dem_set_family_material(ctx, N, mat->load_order, 0), "SetFamilyClumpMaterial");
}
void DEMSolver::SetFamilyMeshMaterial(unsigned int N, const std::shared_ptr<DEMMaterial>& mat) {
    assertInit("SetFamilyMeshMaterial");
    check(dem_set_family_material(ctx, N, mat->load_order, 1), "SetFamilyMeshMaterial");
}
void DEMSolver::SetFamilyPrescribedLinVel(unsigned int ID, const std::string& velX, const std::string& velY,
                                          const std::string& velZ, bool dictate) {
    if (ID > 255) fail("Family numbers must not exceed 255.");
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    const std::string* s[3] = {&velX, &velY, &velZ};
    for (int k = 0; k < 3; k++) {
        float v;
        // (src/DEM/APIPublic.cpp:1027-1047: with dictate, a linear-velocity prescription also holds the rotation -- and a
        // component given a formula is dictated in any case; several prescriptions of one family add up, APIPrivate.cpp:898-908)
        p.linVelP[k] = p.linVelP[k] || dictate;
        p.rotVelP[k] = p.rotVelP[k] || dictate;
        if (parse_prescription(*s[k], simTimeOrZero(), v, p.eLinVel[k])) { p.hasLinVel[k] = true; p.linVel[k] = v; p.linVelP[k] = true; }
    }
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::SetFamilyPrescribedAngVel(unsigned int ID, const std::string& velX, const std::string& velY,
                                          const std::string& velZ, bool dictate) {
    if (ID > 255) fail("Family numbers must not exceed 255.");
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    const std::string* s[3] = {&velX, &velY, &velZ};
    for (int k = 0; k < 3; k++) {
        float v;
        // (src/DEM/APIPublic.cpp:1129-1150: with dictate, an angular-velocity prescription also holds the linear motion)
        p.rotVelP[k] = p.rotVelP[k] || dictate;
        p.linVelP[k] = p.linVelP[k] || dictate;
        if (parse_prescription(*s[k], simTimeOrZero(), v, p.eRotVel[k])) { p.hasRotVel[k] = true; p.rotVel[k] = v; p.rotVelP[k] = true; }
    }
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::SetFamilyPrescribedPosition(unsigned int ID, const std::string& X, const std::string& Y,
                                            const std::string& Z, bool dictate) {
    if (ID > 255) fail("Family numbers must not exceed 255.");
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    const std::string* s[3] = {&X, &Y, &Z};
    for (int k = 0; k < 3; k++) {
        float v;
        // (src/DEM/APIPublic.cpp:1232-1250: with dictate, a position prescription also dictates the orientation)
        p.linPosP[k] = p.linPosP[k] || dictate;
        if (parse_prescription(*s[k], simTimeOrZero(), v, p.eLinPos[k])) { p.hasLinPos[k] = true; p.linPos[k] = v; p.linPosP[k] = true; }
    }
    p.rotPosP = p.rotPosP || dictate;
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::AddFamilyPrescribedAcc(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z) {
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    const std::string* s[3] = {&X, &Y, &Z};
    for (int k = 0; k < 3; k++) {
        float v;
        if (parse_prescription(*s[k], simTimeOrZero(), v, p.eAcc[k])) { p.hasAcc[k] = true; p.acc[k] = v; }
    }
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::AddFamilyPrescribedAngAcc(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z) {
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    const std::string* s[3] = {&X, &Y, &Z};
    for (int k = 0; k < 3; k++) {
        float v;
        if (parse_prescription(*s[k], simTimeOrZero(), v, p.eAngAcc[k])) { p.hasAngAcc[k] = true; p.angAcc[k] = v; }
    }
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::markPrescribed(unsigned int ID, int what, int axes) {
    if (ID > 255) fail("Family numbers must not exceed 255.");
    Prescription& p = m_prescriptions[ID];
    p.used = true;
    for (int k = 0; k < 3; k++) {
        if (!(axes & (1 << k))) continue;
        if (what == 0) p.linVelP[k] = true;
        else if (what == 1) p.rotVelP[k] = true;
        else if (what == 2) p.linPosP[k] = true;
    }
    if (what == 3) p.rotPosP = true;
    if (sys_initialized) uploadFamilies();
}
void DEMSolver::SetFamilyPrescribedQuaternion(unsigned int ID, const std::string& q_formula, bool dictate) {
    std::string t;
    for (char c : q_formula)
        if (!isspace((unsigned char)c)) t.push_back(c);
    if (!t.empty() && t != "none")
        fail("SetFamilyPrescribedQuaternion: a quaternion FORMULA is C++ code the reference compiles at run time; this "
             "ahead-of-time compiled core only supports the dictated form (no formula): set the orientation through a "
             "tracker, or prescribe the angular velocity instead.");
    // (src/DEM/APIPublic.cpp:1328-1332: with dictate, the linear position is dictated as well)
    if (dictate) { markPrescribed(ID, 3, 7); markPrescribed(ID, 2, 7); }
}
void DEMSolver::ChangeFamilyWhen(unsigned int, unsigned int, const std::string&) {
    fail("ChangeFamilyWhen: condition strings need runtime compilation, which this ahead-of-time compiled core does "
         "not have. Use ChangeFamily(ID_from, ID_to) between DoDynamics calls.");
}
void DEMSolver::SetFamilyExtraMargin(unsigned int N, float extra_size) {
    if (N > 255) fail("Family numbers must not exceed 255.");
    m_family_extra_margin[N] = extra_size;
    if (sys_initialized) uploadFamilies();
}

void DEMSolver::uploadFamilies() {
    std::vector<float> extra(DEM_NUM_FAMILIES, 0.f);
    for (const auto& kv : m_family_extra_margin) extra[kv.first] = kv.second;
    std::vector<DemPrescription> pr(DEM_NUM_FAMILIES);
    memset(pr.data(), 0, sizeof(DemPrescription) * DEM_NUM_FAMILIES);
    // the reserved family is always fixed (APIPublic.cpp:980-1011)
    Prescription fixed;
    fixed.used = true;
    for (int k = 0; k < 3; k++) {
        fixed.linVelP[k] = fixed.rotVelP[k] = fixed.linPosP[k] = true;
        fixed.hasLinVel[k] = fixed.hasRotVel[k] = true;
    }
    fixed.rotPosP = true;
    auto put = [&](unsigned fam, const Prescription& p) {
        DemPrescription& d = pr[fam];
        d.used = p.used;
        for (int k = 0; k < 3; k++) {
            d.linVelPrescribed[k] = p.linVelP[k]; d.rotVelPrescribed[k] = p.rotVelP[k]; d.linPosPrescribed[k] = p.linPosP[k];
            d.hasLinVel[k] = p.hasLinVel[k]; d.hasRotVel[k] = p.hasRotVel[k]; d.hasLinPos[k] = p.hasLinPos[k];
            d.hasAcc[k] = p.hasAcc[k]; d.hasAngAcc[k] = p.hasAngAcc[k];
            d.linVel[k] = p.linVel[k]; d.rotVel[k] = p.rotVel[k]; d.linPos[k] = p.linPos[k];
            d.acc[k] = p.acc[k]; d.angAcc[k] = p.angAcc[k];
        }
        d.rotPosPrescribed = p.rotPosP;
    };
    put(RESERVED_FAMILY_NUM, fixed);
    // time-dependent components take their value at the time the next step starts from (the reference hands the
    // integrator simParams->timeElapsed, which it advances after the step: dT.cpp:2463)
    const double t_now = simTimeOrZero();
    for (auto& kv : m_prescriptions) {
        Prescription& p = kv.second;
        for (int k = 0; k < 3; k++) {
            if (p.eLinVel[k]) p.linVel[k] = (float)p.eLinVel[k]->Eval(t_now);
            if (p.eRotVel[k]) p.rotVel[k] = (float)p.eRotVel[k]->Eval(t_now);
            if (p.eLinPos[k]) p.linPos[k] = (float)p.eLinPos[k]->Eval(t_now);
            if (p.eAcc[k]) p.acc[k] = (float)p.eAcc[k]->Eval(t_now);
            if (p.eAngAcc[k]) p.angAcc[k] = (float)p.eAngAcc[k]->Eval(t_now);
        }
    }
    for (const auto& kv : m_prescriptions) put(kv.first, kv.second);
    check(dem_upload_families(ctx, m_family_masks.data(), extra.data(), pr.data()), "dem_upload_families");
}

// ---------------------------------------------------------------------------------------------------------------
void DEMSolver::Initialize(bool dry_run) {
    if (m_loaded_materials.empty()) fail("Before initializing the system, at least one material type should be loaded via LoadMaterial.");
    if (m_ts_size <= 0.0) fail("Time step size is set to be " + std::to_string(m_ts_size) + ". Please supply a positive number via SetInitTimeStep.");

    for (const auto& b : m_cached_input_clump_batches)
        if (!b->owner_wildcards.empty() || !b->geo_wildcards.empty())
            fail("A clump batch carries owner / geometry wildcards (" +
                 (b->owner_wildcards.empty() ? b->geo_wildcards.begin()->first : b->owner_wildcards.begin()->first) +
                 "), but the force model in use declares none. Wildcards of this kind belong to custom force models, "
                 "which need run-time compilation that this ahead-of-time compiled core does not have.");
    for (const auto& m : m_cached_meshes)
        if (!m->geo_wildcards.empty())
            fail("A mesh carries geometry wildcards (" + m->geo_wildcards.begin()->first + "), but the force model in use "
                 "declares none. Wildcards of this kind belong to custom force models, which need run-time compilation "
                 "that this ahead-of-time compiled core does not have.");
    // material properties the force model reads (equipMaterials, APIPrivate.cpp:1882-1933): missing ones default to 0
    if (m_force_model_obj && verbosity >= WARNING)
        for (const std::string& prop_name : m_force_model_obj->m_must_have_mat_props)
            for (const auto& mat : m_loaded_materials)
                if (!mat->mat_prop.count(prop_name)) {
                    std::cerr << "WARNING! Material property " << prop_name << " is needed by the force model or is "
                              << "referred to by the user. However, at least one material does not have it defined, so it "
                              << "is defaulted to 0 for that material.\nPlease be sure this is intentional." << std::endl;
                    break;
                }

    // ---- world sizing (figureOutNV) ----
    DemSimParams sp;
    memset(&sp, 0, sizeof(sp));
    const float tmin[3] = {m_target_box_min.x, m_target_box_min.y, m_target_box_min.z};
    const float tmax[3] = {m_target_box_max.x, m_target_box_max.y, m_target_box_max.z};
    uint32_t nv[3];
    dem_host_figure_out_nv_exact(tmin, tmax, m_box_dir_exact, nv, &sp.l, &sp.voxelSize);
    sp.nvXp2 = nv[0]; sp.nvYp2 = nv[1]; sp.nvZp2 = nv[2];
    sp.integrator = (m_integrator == TIME_INTEGRATOR::FORWARD_EULER) ? DEM_FORWARD_EULER
                    : (m_integrator == TIME_INTEGRATOR::CENTERED_DIFFERENCE) ? DEM_CENTERED_DIFFERENCE : DEM_EXTENDED_TAYLOR;
    sp.force_model = (m_force_model == FORCE_MODEL::HERTZIAN_FRICTIONLESS) ? DEM_HERTZIAN_FRICTIONLESS : DEM_HERTZIAN;
    sp.cd_update_freq = (uint32_t)m_cd_update_freq;
    const float umin[3] = {m_user_box_min.x, m_user_box_min.y, m_user_box_min.z};
    const float umax[3] = {m_user_box_max.x, m_user_box_max.y, m_user_box_max.z};
    const float g[3] = {G.x, G.y, G.z};
    for (int k = 0; k < 3; k++) { sp.LBF[k] = tmin[k]; sp.G[k] = g[k]; sp.userBoxMin[k] = umin[k]; sp.userBoxMax[k] = umax[k]; }
    sp.h = (float)m_ts_size;
    sp.beta = m_expand_factor;
    sp.approxMaxVel = m_approx_max_vel;
    sp.expSafetyMulti = m_expand_safety_multi;
    sp.expSafetyAdder = m_expand_base_vel;
    sp.errOutVel = threshold_error_out_vel;
    sp.record_contact_forces = no_recording_contact_forces ? 0u : 1u;
    check(dem_set_params(ctx, &sp), "dem_set_params");
    m_sp_blob.assign((const unsigned char*)&sp, (const unsigned char*)&sp + sizeof(sp));  // (for UpdateSimParams)

    // ---- world bounding box: one external object appended now (addWorldBoundingBox, APIPrivate.cpp:955-1014) ----
    std::vector<std::shared_ptr<DEMExternObj>> ext = m_cached_extern_objs;
    if (m_user_add_bounding_box != "none") {
        const std::string& m = m_user_add_bounding_box;
        const bool bottom = (m == "only_bottom" || m == "top_open" || m == "all");
        const bool sides = (m == "only_sides" || m == "top_open" || m == "all");
        const bool top = (m == "all");
        auto box = std::make_shared<DEMExternObj>();
        const float3 c = (m_user_box_min + m_user_box_max) / 2.f;
        if (bottom) box->AddPlane(make_float3(c.x, c.y, m_user_box_min.z), make_float3(0, 0, 1), m_bounding_box_material);
        if (sides) {
            box->AddPlane(make_float3(m_user_box_min.x, c.y, c.z), make_float3(1, 0, 0), m_bounding_box_material);
            box->AddPlane(make_float3(m_user_box_max.x, c.y, c.z), make_float3(-1, 0, 0), m_bounding_box_material);
            box->AddPlane(make_float3(c.x, m_user_box_min.y, c.z), make_float3(0, 1, 0), m_bounding_box_material);
            box->AddPlane(make_float3(c.x, m_user_box_max.y, c.z), make_float3(0, -1, 0), m_bounding_box_material);
        }
        if (top) box->AddPlane(make_float3(c.x, c.y, m_user_box_max.z), make_float3(0, 0, -1), m_bounding_box_material);
        ext.push_back(box);
    }

    // ---- templates ----
    std::vector<float> radii, relX, relY, relZ, mass, moiX, moiY, moiZ;
    std::vector<uint32_t> comp_start;
    std::vector<uint16_t> comp_mat;
    for (const auto& t : m_templates) {
        comp_start.push_back((uint32_t)radii.size());
        for (size_t k = 0; k < t->radii.size(); k++) {
            radii.push_back(t->radii[k]);
            relX.push_back(t->relPos[k].x); relY.push_back(t->relPos[k].y); relZ.push_back(t->relPos[k].z);
            comp_mat.push_back((uint16_t)t->materials[k]->load_order);
        }
        mass.push_back(t->mass); moiX.push_back(t->MOI.x); moiY.push_back(t->MOI.y); moiZ.push_back(t->MOI.z);
    }
    for (const auto& e : ext) { mass.push_back(e->mass); moiX.push_back(e->MOI.x); moiY.push_back(e->MOI.y); moiZ.push_back(e->MOI.z); }
    for (const auto& m : m_cached_meshes) { mass.push_back(m->mass); moiX.push_back(m->MOI.x); moiY.push_back(m->MOI.y); moiZ.push_back(m->MOI.z); }
    check(dem_upload_templates(ctx, (uint32_t)radii.size(), radii.data(), relX.data(), relY.data(), relZ.data(),
                               (uint32_t)mass.size(), mass.data(), moiX.data(), moiY.data(), moiZ.data()),
          "dem_upload_templates");

    // ---- materials (equipMaterials, APIPrivate.cpp:1877-2026) ----
    const uint32_t nM = (uint32_t)m_loaded_materials.size();
    std::vector<float> E(nM), nu(nM), CoR(nM * nM), mu(nM * nM), Crr(nM * nM);
    auto prop = [&](uint32_t i, const char* name) {
        const auto& mp = m_loaded_materials[i]->mat_prop;
        auto it = mp.find(name);
        return it == mp.end() ? 0.f : it->second;
    };
    for (uint32_t i = 0; i < nM; i++) { E[i] = prop(i, "E"); nu[i] = prop(i, "nu"); }
    struct { const char* name; std::vector<float>* t; } pw[3] = {{"CoR", &CoR}, {"mu", &mu}, {"Crr", &Crr}};
    for (auto& p : pw) {
        for (uint32_t i = 0; i < nM; i++) (*p.t)[i * nM + i] = prop(i, p.name);
        for (uint32_t i = 0; i < nM; i++)
            for (uint32_t j = 0; j < nM; j++)
                if (i != j) (*p.t)[i * nM + j] = (float)(((*p.t)[i * nM + i] + (*p.t)[j * nM + j]) / 2.);
        auto it = m_pairwise_matprop.find(p.name);
        if (it != m_pairwise_matprop.end())
            for (const auto& kv : it->second) {
                (*p.t)[kv.first.first * nM + kv.first.second] = kv.second;
                (*p.t)[kv.first.second * nM + kv.first.first] = kv.second;
            }
    }
    check(dem_upload_materials(ctx, nM, E.data(), nu.data(), CoR.data(), mu.data(), Crr.data()), "dem_upload_materials");
    m_any_rolling_resistance = std::any_of(Crr.begin(), Crr.end(), [](float c) { return c > 0.f; });

    // ---- owners: clumps, then external objects, then meshes (dT.cpp:638-1024) ----
    size_t nC = 0;
    for (const auto& b : m_cached_input_clump_batches) nC += b->nClumps;
    const size_t nE = ext.size(), nMesh = m_cached_meshes.size(), nO = nC + nE + nMesh;
    std::vector<float> xyz(3 * nO), qw(nO), qx(nO), qy(nO), qz(nO), vx(nO), vy(nO), vz(nO), ox(nO), oy(nO), oz(nO);
    std::vector<uint8_t> fam(nO);
    std::vector<uint16_t> inertia(nO);
    std::vector<uint32_t> sph_owner;
    std::vector<uint16_t> sph_comp, sph_mat;
    m_owner_mass.assign(nO, 0.f);
    m_owner_moi.assign(nO, make_float3(0, 0, 0));
    m_owner_type_mark.assign(nC, 0);
    size_t o = 0;
    bool out_of_box = false;
    for (auto& b : m_cached_input_clump_batches) {
        b->first_owner = (bodyID_t)o;
        for (size_t j = 0; j < b->nClumps; j++, o++) {
            const auto& t = b->types[j];
            xyz[3 * o] = b->xyz[j].x; xyz[3 * o + 1] = b->xyz[j].y; xyz[3 * o + 2] = b->xyz[j].z;
            if (b->xyz[j].x < m_user_box_min.x || b->xyz[j].x > m_user_box_max.x || b->xyz[j].y < m_user_box_min.y ||
                b->xyz[j].y > m_user_box_max.y || b->xyz[j].z < m_user_box_min.z || b->xyz[j].z > m_user_box_max.z)
                out_of_box = true;
            qw[o] = b->oriQ[j].w; qx[o] = b->oriQ[j].x; qy[o] = b->oriQ[j].y; qz[o] = b->oriQ[j].z;
            vx[o] = b->vel[j].x; vy[o] = b->vel[j].y; vz[o] = b->vel[j].z;
            ox[o] = b->angVel[j].x; oy[o] = b->angVel[j].y; oz[o] = b->angVel[j].z;
            fam[o] = (uint8_t)b->families[j];
            inertia[o] = (uint16_t)t->mark;
            m_owner_mass[o] = t->mass; m_owner_moi[o] = t->MOI; m_owner_type_mark[o] = t->mark;
            for (unsigned int k = 0; k < t->nComp; k++) {
                sph_owner.push_back((uint32_t)o);
                sph_comp.push_back((uint16_t)(comp_start[t->mark] + k));
                sph_mat.push_back(comp_mat[comp_start[t->mark] + k]);
            }
        }
    }
    if (out_of_box && verbosity >= WARNING)
        std::cerr << "WARNING! At least one clump is initialized outside the user-specified box domain." << std::endl;
    std::vector<uint32_t> objOwner;
    std::vector<uint8_t> objType;
    std::vector<uint16_t> objMat;
    std::vector<float> objNormal, rpx, rpy, rpz, rtx, rty, rtz, s1, s2, s3, objMass;
    for (size_t e = 0; e < nE; e++, o++) {
        auto& ob = ext[e];
        ob->owner = (bodyID_t)o;
        xyz[3 * o] = ob->init_pos.x; xyz[3 * o + 1] = ob->init_pos.y; xyz[3 * o + 2] = ob->init_pos.z;
        qw[o] = ob->init_oriQ.w; qx[o] = ob->init_oriQ.x; qy[o] = ob->init_oriQ.y; qz[o] = ob->init_oriQ.z;
        fam[o] = (uint8_t)ob->family_code;
        inertia[o] = (uint16_t)(m_templates.size() + e);
        m_owner_mass[o] = ob->mass; m_owner_moi[o] = ob->MOI;
        for (const auto& c : ob->comps) {
            objOwner.push_back((uint32_t)o); objType.push_back((uint8_t)c.type);
            objMat.push_back((uint16_t)c.material->load_order); objNormal.push_back(c.normal);
            rpx.push_back(c.pos.x); rpy.push_back(c.pos.y); rpz.push_back(c.pos.z);
            rtx.push_back(c.dir.x); rty.push_back(c.dir.y); rtz.push_back(c.dir.z);
            s1.push_back(c.size1); s2.push_back(0.f); s3.push_back(0.f); objMass.push_back(ob->mass);
        }
    }
    std::vector<uint32_t> triOwner;
    std::vector<uint16_t> triMat;
    std::vector<float> tn1, tn2, tn3;
    for (size_t mi = 0; mi < nMesh; mi++, o++) {
        auto& me = m_cached_meshes[mi];
        if (me->nTri > 0 && !me->isMaterialSet)
            fail("A meshed object is loaded but does not have associated material.\nPlease assign material to meshes via SetMaterial.");
        me->owner = (bodyID_t)o;
        me->tri_first = triOwner.size();
        xyz[3 * o] = me->init_pos.x; xyz[3 * o + 1] = me->init_pos.y; xyz[3 * o + 2] = me->init_pos.z;
        qw[o] = me->init_oriQ.w; qx[o] = me->init_oriQ.x; qy[o] = me->init_oriQ.y; qz[o] = me->init_oriQ.z;
        fam[o] = (uint8_t)me->family_code;
        inertia[o] = (uint16_t)(m_templates.size() + nE + mi);
        m_owner_mass[o] = me->mass; m_owner_moi[o] = me->MOI;
        for (size_t t = 0; t < me->nTri; t++) {
            const int3 fc = me->m_face_v_indices[t];
            const float3 a = me->m_vertices[fc.x], b = me->m_vertices[fc.y], c = me->m_vertices[fc.z];
            triOwner.push_back((uint32_t)o);
            triMat.push_back((uint16_t)me->materials[t]->load_order);
            tn1.insert(tn1.end(), {a.x, a.y, a.z});
            tn2.insert(tn2.end(), {b.x, b.y, b.z});
            tn3.insert(tn3.end(), {c.x, c.y, c.z});
        }
    }
    std::vector<uint64_t> voxel(nO);
    std::vector<uint16_t> lx(nO), ly(nO), lz(nO);
    dem_host_encode_positions(&sp, xyz.data(), nO, voxel.data(), lx.data(), ly.data(), lz.data());
    if (m_carry.nBodies) {
        // UpdateClumps: the owners that were already in the simulation keep their exact state (position codes,
        // orientation, velocities, family). Old clumps keep their indices; the external objects and meshes move up
        // behind the newly added clumps.
        if (nE + nMesh != m_carry.nBodies - m_carry.nClumps || nC < m_carry.nClumps)
            fail("UpdateClumps: only clumps can be added to an initialised system.");
        for (size_t i = 0; i < m_carry.nBodies; i++) {
            const size_t j = (i < m_carry.nClumps) ? i : i - m_carry.nClumps + nC;
            voxel[j] = m_carry.voxel[i]; lx[j] = m_carry.lx[i]; ly[j] = m_carry.ly[i]; lz[j] = m_carry.lz[i];
            qw[j] = m_carry.quat[4 * i]; qx[j] = m_carry.quat[4 * i + 1]; qy[j] = m_carry.quat[4 * i + 2]; qz[j] = m_carry.quat[4 * i + 3];
            vx[j] = m_carry.vel[3 * i]; vy[j] = m_carry.vel[3 * i + 1]; vz[j] = m_carry.vel[3 * i + 2];
            ox[j] = m_carry.omg[3 * i]; oy[j] = m_carry.omg[3 * i + 1]; oz[j] = m_carry.omg[3 * i + 2];
            fam[j] = m_carry.fam[i];
        }
    }
    check(dem_upload_analytical(ctx, (uint32_t)objOwner.size(), objOwner.data(), objType.data(), objMat.data(),
                                objNormal.data(), rpx.data(), rpy.data(), rpz.data(), rtx.data(), rty.data(), rtz.data(),
                                s1.data(), s2.data(), s3.data(), objMass.data()),
          "dem_upload_analytical");
    uploadFamilies();
    check(dem_upload_owners(ctx, (uint32_t)nO, voxel.data(), lx.data(), ly.data(), lz.data(), qw.data(), qx.data(),
                            qy.data(), qz.data(), vx.data(), vy.data(), vz.data(), ox.data(), oy.data(), oz.data(),
                            fam.data(), inertia.data()),
          "dem_upload_owners");
    check(dem_upload_spheres(ctx, (uint32_t)sph_owner.size(), sph_owner.data(), sph_comp.data(), sph_mat.data()),
          "dem_upload_spheres");
    check(dem_upload_triangles(ctx, (uint32_t)triOwner.size(), triOwner.data(), tn1.data(), tn2.data(), tn3.data(), triMat.data()),
          "dem_upload_triangles");
    check(dem_initialize(ctx, 0), "dem_initialize");
    check(dem_set_option(ctx, "update_freq_max", (double)m_max_update_freq), "SetCDMaxUpdateFreq");
    check(dem_set_option(ctx, "adaptive_update_freq", m_adaptive_update_freq ? 1.0 : 0.0), "UseAdaptiveUpdateFreq");
    nOwnerClumps = nC; nOwnerBodies = nO; nSpheres = sph_owner.size();
    m_sphere_owner = sph_owner;
    m_owner_first_sphere.assign(nO + 1, (unsigned int)sph_owner.size());  // spheres are numbered owner by owner
    for (size_t s_id = sph_owner.size(); s_id-- > 0;) m_owner_first_sphere[sph_owner[s_id]] = (unsigned int)s_id;
    m_tri_owner = triOwner;
    m_anal_owner = objOwner;
    if (!m_trackers.empty() || (m_out_content & (ABS_ACC | ACC | ANG_ACC)))
        check(dem_set_option(ctx, "keep_acc", 1.0), "dem_set_option");

    // ---- restart: existing contacts + wildcards of the batches (Structs.h:857-882, dT.cpp:849-881) ----
    {
        std::vector<uint32_t> idA, idB;
        std::vector<uint8_t> type;
        std::vector<float> wc;
        size_t sphere_base = 0, batch_no = 0;
        if (m_carry.nBodies) {  // UpdateClumps: the contacts (and their history) of the running simulation
            idA = m_carry.idA; idB = m_carry.idB; type = m_carry.ctype; wc = m_carry.wildcards;
        }
        for (const auto& b : m_cached_input_clump_batches) {
            // (the contact pairs of batches that were already initialised went in then and have evolved since)
            const size_t n = (batch_no++ < m_carry.nBatches) ? 0 : b->contact_pairs.size();
            const char* names[4] = {"delta_tan_x", "delta_tan_y", "delta_tan_z", "delta_time"};
            for (size_t i = 0; i < n; i++) {
                idA.push_back((uint32_t)(sphere_base + b->contact_pairs[i].first));
                idB.push_back((uint32_t)(sphere_base + b->contact_pairs[i].second));
                type.push_back(DEM_CNT_SPHERE_SPHERE);
                for (int k = 0; k < 4; k++) {
                    auto it = b->contact_wildcards.find(names[k]);
                    wc.push_back(it != b->contact_wildcards.end() && it->second.size() == n ? it->second[i] : 0.f);
                }
            }
            sphere_base += b->nSpheres;
        }
        if (!idA.empty()) check(dem_set_contacts(ctx, idA.size(), idA.data(), idB.data(), type.data(), wc.data()), "dem_set_contacts");
    }
    sys_initialized = true;
    m_n_init_batches = m_cached_input_clump_batches.size();
    if (verbosity >= INFO)
        std::cout << "DEM core initialised: " << nC << " clumps, " << nSpheres << " spheres, " << objOwner.size()
                  << " analytical components; l = " << sp.l << ", voxel bits " << sp.nvXp2 << "/" << sp.nvYp2 << "/"
                  << sp.nvZp2 << std::endl;
    // the reference's Initialize ends with a dry run that builds the first contact list (APIPublic.cpp:2207-2212)
    if (dry_run) DoDynamicsThenSync(0.0);
}

// UpdateClumps (APIPublic.cpp:2347-2390): clumps added with AddClumps after Initialize() join the running simulation.
// The flattened arrays are rebuilt and uploaded again; owners that were already there keep their exact device state and
// the contact list keeps its history (dem_set_contacts). New clumps are numbered right behind the existing clumps (the
// reference numbers them behind ALL existing owners; objects are addressed through their handles here, so only the raw
// owner ids of external objects / meshes differ).
void DEMSolver::UpdateClumps() {
    assertInit("UpdateClumps");
    if (m_cached_input_clump_batches.size() == m_n_init_batches) {
        if (verbosity >= WARNING) std::cerr << "WARNING! UpdateClumps is called, but no new clumps were added since the last initialization." << std::endl;
        return;
    }
    CarriedState& c = m_carry;
    c.nClumps = nOwnerClumps; c.nBodies = nOwnerBodies; c.nBatches = m_n_init_batches;
    const size_t n = nOwnerBodies;
    c.voxel.resize(n); c.lx.resize(n); c.ly.resize(n); c.lz.resize(n);
    c.quat.resize(4 * n); c.vel.resize(3 * n); c.omg.resize(3 * n); c.fam.resize(n);
    check(dem_download_owner_state(ctx, 0, (uint32_t)n, c.voxel.data(), c.lx.data(), c.ly.data(), c.lz.data(), c.quat.data(),
                                   c.vel.data(), c.omg.data(), nullptr, nullptr, c.fam.data()), "dem_download_owner_state");
    uint64_t nc = 0;
    check(dem_download_contacts(ctx, 0, &nc, nullptr, nullptr, nullptr, nullptr, nullptr), "dem_download_contacts");
    c.idA.resize(nc); c.idB.resize(nc); c.ctype.resize(nc); c.wildcards.resize(4 * nc);
    if (nc) check(dem_download_contacts(ctx, nc, &nc, c.idA.data(), c.idB.data(), c.ctype.data(), c.wildcards.data(), nullptr), "dem_download_contacts");
    Initialize(false);
    c = CarriedState();
    DoDynamicsThenSync(0.0);  // the contact list now includes the newcomers
}
void DEMSolver::ClearCache() {}

bool DEMSolver::anyTimeDependentPrescription() const {
    for (const auto& kv : m_prescriptions)
        if (kv.second.TimeDependent()) return true;
    return false;
}
double DEMSolver::simTimeOrZero() const { return sys_initialized ? GetSimTime() : 0.0; }

void DEMSolver::DoDynamics(double thisCallDuration) {
    assertInit("DoDynamics");
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (thisCallDuration > 0.0 && anyTimeDependentPrescription()) {
        // prescriptions that depend on t: refresh the family tables before every step (one stream-ordered 56 KB copy,
        // no synchronisation), same step count as the reference's loop (dT.cpp:2401)
        const double h = (double)(float)m_ts_size;
        uint64_t n = 0;
        for (double cycle = 0.0; cycle < thisCallDuration; cycle += h) n++;
        for (uint64_t i = 0; i < n; i++) {
            uploadFamilies();
            check(dem_step_async(ctx, 1), "DoDynamics");
        }
        check(dem_sync(ctx), "DoDynamics");
    } else
    check(dem_do_dynamics(ctx, thisCallDuration), "DoDynamics");
    m_wall_time_dynamics += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    // the reference's contact detection stops a run whose lists explode (DEMCubContactDetection.cu:876-892)
    const float avg = GetAvgSphContacts();
    if (avg > threshold_error_out_num_cnts)
        fail("On average a sphere has " + std::to_string(avg) + " contacts, more than the max allowance (" +
             std::to_string(threshold_error_out_num_cnts) + ").\nIf you believe this is not abnormal, set the allowance high "
             "using SetErrorOutAvgContacts before initialization.\nIf you think this is because the contact margin added is too "
             "big, use SetCDMaxUpdateFreq to limit the number of steps a contact list is used for.\nOtherwise, the simulation "
             "may have diverged and relaxing the physics may help, such as decreasing the step size and modifying material "
             "properties.\nIf this happens at the start of simulation, check if there are initial penetrations, a.k.a. elements "
             "initialized inside walls.");
}
void DEMSolver::DoDynamicsThenSync(double thisCallDuration) { DoDynamics(thisCallDuration); }

void DEMSolver::ChangeFamily(unsigned int ID_from, unsigned int ID_to) {
    assertInit("ChangeFamily");
    if (ID_from > 255 || ID_to > 255) fail("Family numbers must not exceed 255.");
    std::vector<uint8_t> f(nOwnerBodies);
    check(dem_download_owner_state(ctx, 0, (uint32_t)nOwnerBodies, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, f.data()), "dem_download_owner_state");
    for (auto& x : f)
        if (x == ID_from) x = (uint8_t)ID_to;
    check(dem_upload_owner_state(ctx, 0, (uint32_t)nOwnerBodies, nullptr, nullptr, nullptr, nullptr, f.data()), "dem_upload_owner_state");
}

size_t DEMSolver::ChangeClumpFamily(unsigned int fam_num, const std::pair<double, double>& X, const std::pair<double, double>& Y,
                                    const std::pair<double, double>& Z, const std::set<unsigned int>& orig_fam) {
    assertInit("ChangeClumpFamily");
    if (fam_num > 255) fail("Family numbers must not exceed 255.");
    const uint32_t n = (uint32_t)nOwnerClumps;
    std::vector<float> pos(3 * (size_t)n);
    std::vector<uint8_t> f(n);
    check(dem_download_positions(ctx, 0, n, pos.data(), nullptr), "dem_download_positions");
    check(dem_download_owner_state(ctx, 0, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   f.data()), "dem_download_owner_state");
    size_t changed = 0;
    for (uint32_t i = 0; i < n; i++) {
        const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        if (x < X.first || x > X.second || y < Y.first || y > Y.second || z < Z.first || z > Z.second) continue;
        if (!orig_fam.empty() && !orig_fam.count(f[i])) continue;
        if (f[i] != (uint8_t)fam_num) changed++;
        f[i] = (uint8_t)fam_num;
    }
    if (changed) check(dem_upload_owner_state(ctx, 0, n, nullptr, nullptr, nullptr, nullptr, f.data()), "dem_upload_owner_state");
    return changed;
}

size_t DEMSolver::GetNumContacts() const {
    DemStats s;
    dem_get_stats(ctx, &s);
    return (size_t)(s.n_contacts_ss + s.n_contacts_sa + s.n_contacts_st);
}
float DEMSolver::GetAvgSphContacts() const {
    DemStats s;
    dem_get_stats(ctx, &s);
    return nSpheres ? (float)((double)(s.n_contacts_ss + s.n_contacts_sa + s.n_contacts_st) / (double)nSpheres) : 0.f;
}
void DEMSolver::SetSimTime(double time) { check(dem_set_sim_time(ctx, time), "dem_set_sim_time"); }
double DEMSolver::GetBinSize() const {
    DemStats st;
    check(dem_get_stats(ctx, &st), "dem_get_stats");
    return st.cell_size;
}
size_t DEMSolver::GetBinNum() const {
    DemStats st;
    check(dem_get_stats(ctx, &st), "dem_get_stats");
    return (size_t)st.n_cells[0] * st.n_cells[1] * st.n_cells[2];
}
float DEMSolver::GetExpandFactor() const {
    DemStats st;
    check(dem_get_stats(ctx, &st), "dem_get_stats");
    return st.max_margin;
}
size_t DEMSolver::GetDeviceMemUsageDynamic() const {
    DemStats st;
    check(dem_get_stats(ctx, &st), "dem_get_stats");
    return (size_t)st.device_bytes;
}
std::vector<bodyID_t> DEMSolver::GetOwnerContactClumps(bodyID_t ownerID) const {
    std::vector<bodyID_t> out;
    for (const auto& pr : GetClumpContacts()) {
        if (pr.first == ownerID) out.push_back(pr.second);
        else if (pr.second == ownerID) out.push_back(pr.first);
    }
    return out;
}
double DEMSolver::GetSimTime() const {
    DemStats s;
    dem_get_stats(ctx, &s);
    return s.sim_time;
}
void DEMSolver::ShowThreadCollaborationStats() {
    DemStats s;
    dem_get_stats(ctx, &s);
    std::cout << "Number of steps: " << s.n_steps << "\nNumber of contact-list rebuilds: " << s.n_rebuilds
              << "\nAverage steps per rebuild: " << (s.n_rebuilds ? (double)s.n_steps / (double)s.n_rebuilds : 0.0)
              << "\n(one in-order CUDA stream: no kT/dT hand-shake, no held-back steps)" << std::endl;
}
void DEMSolver::ShowTimingStats() {
    DemStats s;
    dem_get_stats(ctx, &s);
    std::cout << "Wall time inside DoDynamics: " << m_wall_time_dynamics << " s for " << s.n_steps << " steps ("
              << (m_wall_time_dynamics > 0 ? s.n_steps / m_wall_time_dynamics : 0.0) << " steps/s), " << s.kernel_launches
              << " kernel launches" << std::endl;
}
void DEMSolver::ShowMemStats() const {
    DemStats s;
    dem_get_stats(ctx, &s);
    std::cout << "Device memory held by the DEM core: " << s.device_bytes / (1024.0 * 1024.0) << " MiB" << std::endl;
}

// ---- raw owner access: n consecutive owners per call (src/DEM/API.h:431-471, dT.cpp:3062-3130 of the reference) ----
namespace {
std::vector<float3> to_float3(const std::vector<float>& v) {
    std::vector<float3> out(v.size() / 3);
    for (size_t i = 0; i < out.size(); i++) out[i] = make_float3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    return out;
}
}  // namespace
#define OWNER_RANGE(what)                                                                       \
    assertInit(what);                                                                           \
    if ((size_t)ownerID + n > nOwnerBodies || n == 0) fail(std::string(what) + ": owner range out of bounds")

std::vector<float3> DEMSolver::GetOwnerPosition(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerPosition");
    std::vector<float> p(3 * (size_t)n);
    check(dem_download_positions(ctx, ownerID, n, p.data(), nullptr), "dem_download_positions");
    return to_float3(p);
}
std::vector<float3> DEMSolver::GetOwnerVelocity(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerVelocity");
    std::vector<float> v(3 * (size_t)n);
    check(dem_download_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, nullptr, v.data(), nullptr, nullptr, nullptr, nullptr), "dem_download_owner_state");
    return to_float3(v);
}
std::vector<float3> DEMSolver::GetOwnerAngVel(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerAngVel");
    std::vector<float> v(3 * (size_t)n);
    check(dem_download_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, v.data(), nullptr, nullptr, nullptr), "dem_download_owner_state");
    return to_float3(v);
}
std::vector<float4> DEMSolver::GetOwnerOriQ(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerOriQ");
    std::vector<float> q(4 * (size_t)n);
    check(dem_download_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, q.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "dem_download_owner_state");
    std::vector<float4> out(n);
    for (size_t i = 0; i < out.size(); i++)  // core stores w,x,y,z; the API speaks x,y,z,w
        out[i] = make_float4(q[4 * i + 1], q[4 * i + 2], q[4 * i + 3], q[4 * i]);
    return out;
}
std::vector<float3> DEMSolver::GetOwnerAcc(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerAcc");
    std::vector<float> v(3 * (size_t)n);
    check(dem_download_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, v.data(), nullptr, nullptr), "dem_download_owner_state");
    return to_float3(v);
}
std::vector<float3> DEMSolver::GetOwnerAngAcc(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerAngAcc");
    std::vector<float> v(3 * (size_t)n);
    check(dem_download_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, v.data(), nullptr), "dem_download_owner_state");
    return to_float3(v);
}
std::vector<unsigned int> DEMSolver::GetOwnerFamily(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerFamily");
    std::vector<uint8_t> f(n);
    check(dem_download_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, f.data()), "dem_download_owner_state");
    return std::vector<unsigned int>(f.begin(), f.end());
}
std::vector<float> DEMSolver::GetOwnerMass(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerMass");
    return std::vector<float>(m_owner_mass.begin() + ownerID, m_owner_mass.begin() + ownerID + n);
}
std::vector<float3> DEMSolver::GetOwnerMOI(bodyID_t ownerID, bodyID_t n) const {
    OWNER_RANGE("GetOwnerMOI");
    return std::vector<float3>(m_owner_moi.begin() + ownerID, m_owner_moi.begin() + ownerID + n);
}
namespace {
std::vector<float> flat3(const std::vector<float3>& v) {
    std::vector<float> out(3 * v.size());
    for (size_t i = 0; i < v.size(); i++) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
    return out;
}
}  // namespace
void DEMSolver::SetOwnerPosition(bodyID_t ownerID, const std::vector<float3>& pos) {
    assertInit("SetOwnerPosition");
    if (pos.empty()) return;
    check(dem_upload_owner_state(ctx, ownerID, (uint32_t)pos.size(), flat3(pos).data(), nullptr, nullptr, nullptr, nullptr), "dem_upload_owner_state");
}
void DEMSolver::SetOwnerVelocity(bodyID_t ownerID, const std::vector<float3>& vel) {
    assertInit("SetOwnerVelocity");
    if (vel.empty()) return;
    check(dem_upload_owner_state(ctx, ownerID, (uint32_t)vel.size(), nullptr, nullptr, flat3(vel).data(), nullptr, nullptr), "dem_upload_owner_state");
}
void DEMSolver::SetOwnerAngVel(bodyID_t ownerID, const std::vector<float3>& angVel) {
    assertInit("SetOwnerAngVel");
    if (angVel.empty()) return;
    check(dem_upload_owner_state(ctx, ownerID, (uint32_t)angVel.size(), nullptr, nullptr, nullptr, flat3(angVel).data(), nullptr), "dem_upload_owner_state");
}
void DEMSolver::SetOwnerOriQ(bodyID_t ownerID, const std::vector<float4>& oriQ) {
    assertInit("SetOwnerOriQ");
    if (oriQ.empty()) return;
    std::vector<float> q(4 * oriQ.size());
    for (size_t i = 0; i < oriQ.size(); i++) { q[4 * i] = oriQ[i].w; q[4 * i + 1] = oriQ[i].x; q[4 * i + 2] = oriQ[i].y; q[4 * i + 3] = oriQ[i].z; }
    check(dem_upload_owner_state(ctx, ownerID, (uint32_t)oriQ.size(), nullptr, q.data(), nullptr, nullptr, nullptr), "dem_upload_owner_state");
}
void DEMSolver::SetOwnerFamily(bodyID_t ownerID, unsigned int fam, bodyID_t n) {
    assertInit("SetOwnerFamily");
    if (fam > 255) fail("SetOwnerFamily: family number " + std::to_string(fam) + " is larger than the max allowance 255");
    if (n == 0) return;
    const std::vector<uint8_t> f(n, (uint8_t)fam);
    check(dem_upload_owner_state(ctx, ownerID, n, nullptr, nullptr, nullptr, nullptr, f.data()), "dem_upload_owner_state");
}
void DEMSolver::AddOwnerNextStepAcc(bodyID_t ownerID, const std::vector<float3>& acc) {
    assertInit("AddOwnerNextStepAcc");
    if (acc.empty()) return;
    check(dem_add_owner_acc(ctx, ownerID, (uint32_t)acc.size(), flat3(acc).data(), nullptr), "dem_add_owner_acc");
}
void DEMSolver::AddOwnerNextStepAngAcc(bodyID_t ownerID, const std::vector<float3>& angAcc) {
    assertInit("AddOwnerNextStepAngAcc");
    if (angAcc.empty()) return;
    check(dem_add_owner_acc(ctx, ownerID, (uint32_t)angAcc.size(), nullptr, flat3(angAcc).data()), "dem_add_owner_acc");
}
void DEMSolver::ChangeClumpSizes(const std::vector<bodyID_t>&, const std::vector<float>&) {
    // (the reference needs flattened, non-jitified templates for this, APIPublic.cpp:2416-2442; this core always
    // addresses sphere components through their clump template, so a single clump cannot be resized)
    fail("ChangeClumpSizes is not available: clump components are stored per template, not per clump. Load a scaled "
         "template (DEMClumpTemplate::Scale) and add the clumps with it instead.");
}
double DEMSolver::Reduce(int kind) const {
    assertInit("inspector");
    double out = 0;
    check(dem_reduce(ctx, kind, &out), "dem_reduce");
    return out;
}

// ---- trackers ----
bodyID_t DEMTracker::first() {
    if (obj->obj_type == OWNER_TYPE::CLUMP) return std::static_pointer_cast<DEMClumpBatch>(obj)->first_owner;
    if (obj->obj_type == OWNER_TYPE::MESH) return std::static_pointer_cast<DEMMeshConnected>(obj)->owner;
    return std::static_pointer_cast<DEMExternObj>(obj)->owner;
}
size_t DEMTracker::count() {
    if (obj->obj_type == OWNER_TYPE::CLUMP) return std::static_pointer_cast<DEMClumpBatch>(obj)->nClumps;
    return 1;
}
bodyID_t DEMTracker::GetOwnerID(size_t offset) {
    if (offset >= count()) fail("Tracker offset exceeds the number of owners it tracks.");
    return first() + (bodyID_t)offset;
}
void DEMTracker::assertOwnerSize(size_t input_length, const std::string& name) {
    if (input_length != count())
        fail(name + " is called with " + std::to_string(input_length) + " values, but this tracker tracks " +
             std::to_string(count()) + " owners.");
}
std::vector<bodyID_t> DEMTracker::GetOwnerIDs() {
    std::vector<bodyID_t> ids(count());
    for (size_t i = 0; i < ids.size(); i++) ids[i] = first() + (bodyID_t)i;
    return ids;
}
float3 DEMTracker::Pos(size_t offset) { return sys->GetOwnerPosition(GetOwnerID(offset))[0]; }
float3 DEMTracker::Vel(size_t offset) { return sys->GetOwnerVelocity(GetOwnerID(offset))[0]; }
float3 DEMTracker::AngVelLocal(size_t offset) { return sys->GetOwnerAngVel(GetOwnerID(offset))[0]; }
float3 DEMTracker::AngVelGlobal(size_t offset) {
    const bodyID_t id = GetOwnerID(offset);
    return Rotate(sys->GetOwnerAngVel(id)[0], sys->GetOwnerOriQ(id)[0]);
}
float4 DEMTracker::OriQ(size_t offset) { return sys->GetOwnerOriQ(GetOwnerID(offset))[0]; }
float3 DEMTracker::ContactAcc(size_t offset) { return sys->GetOwnerAcc(GetOwnerID(offset))[0]; }
float3 DEMTracker::ContactAngAccLocal(size_t offset) { return sys->GetOwnerAngAcc(GetOwnerID(offset))[0]; }
float3 DEMTracker::ContactAngAccGlobal(size_t offset) {
    const bodyID_t id = GetOwnerID(offset);
    return Rotate(sys->GetOwnerAngAcc(id)[0], sys->GetOwnerOriQ(id)[0]);
}
float DEMTracker::Mass(size_t offset) { return sys->GetOwnerMass(GetOwnerID(offset))[0]; }
float3 DEMTracker::MOI(size_t offset) { return sys->GetOwnerMOI(GetOwnerID(offset))[0]; }
unsigned int DEMTracker::GetFamily(size_t offset) { return sys->GetOwnerFamily(GetOwnerID(offset))[0]; }
// the plural forms: one transfer for the whole tracked object
std::vector<float3> DEMTracker::Positions() { return sys->GetOwnerPosition(first(), (bodyID_t)count()); }
std::vector<float3> DEMTracker::Velocities() { return sys->GetOwnerVelocity(first(), (bodyID_t)count()); }
std::vector<float3> DEMTracker::AngularVelocitiesLocal() { return sys->GetOwnerAngVel(first(), (bodyID_t)count()); }
namespace {
std::vector<float3> to_global(std::vector<float3> v, const std::vector<float4>& q) {
    for (size_t i = 0; i < v.size(); i++) v[i] = Rotate(v[i], q[i]);
    return v;
}
}  // namespace
std::vector<float3> DEMTracker::AngularVelocitiesGlobal() {
    return to_global(sys->GetOwnerAngVel(first(), (bodyID_t)count()), sys->GetOwnerOriQ(first(), (bodyID_t)count()));
}
std::vector<float4> DEMTracker::OrientationQuaternions() { return sys->GetOwnerOriQ(first(), (bodyID_t)count()); }
std::vector<unsigned int> DEMTracker::GetFamilies() { return sys->GetOwnerFamily(first(), (bodyID_t)count()); }
std::vector<bodyID_t> DEMTracker::GetContactClumps(size_t offset) { return sys->GetOwnerContactClumps(GetOwnerID(offset)); }
std::vector<float3> DEMTracker::ContactAccelerations() { return sys->GetOwnerAcc(first(), (bodyID_t)count()); }
std::vector<float3> DEMTracker::ContactAngularAccelerationsLocal() { return sys->GetOwnerAngAcc(first(), (bodyID_t)count()); }
std::vector<float3> DEMTracker::ContactAngularAccelerationsGlobal() {
    return to_global(sys->GetOwnerAngAcc(first(), (bodyID_t)count()), sys->GetOwnerOriQ(first(), (bodyID_t)count()));
}
std::vector<float> DEMTracker::Masses() { return sys->GetOwnerMass(first(), (bodyID_t)count()); }
std::vector<float3> DEMTracker::MOIs() { return sys->GetOwnerMOI(first(), (bodyID_t)count()); }
void DEMTracker::SetPos(float3 pos, size_t offset) { sys->SetOwnerPosition(GetOwnerID(offset), std::vector<float3>{pos}); }
void DEMTracker::SetVel(float3 vel, size_t offset) { sys->SetOwnerVelocity(GetOwnerID(offset), std::vector<float3>{vel}); }
void DEMTracker::SetAngVel(float3 angVel, size_t offset) { sys->SetOwnerAngVel(GetOwnerID(offset), std::vector<float3>{angVel}); }
void DEMTracker::SetOriQ(float4 oriQ, size_t offset) { sys->SetOwnerOriQ(GetOwnerID(offset), std::vector<float4>{oriQ}); }
void DEMTracker::SetPos(const std::vector<float3>& pos) { assertOwnerSize(pos.size(), "SetPos"); sys->SetOwnerPosition(first(), pos); }
void DEMTracker::SetVel(const std::vector<float3>& vel) { assertOwnerSize(vel.size(), "SetVel"); sys->SetOwnerVelocity(first(), vel); }
void DEMTracker::SetAngVel(const std::vector<float3>& angVel) { assertOwnerSize(angVel.size(), "SetAngVel"); sys->SetOwnerAngVel(first(), angVel); }
void DEMTracker::SetOriQ(const std::vector<float4>& oriQ) { assertOwnerSize(oriQ.size(), "SetOriQ"); sys->SetOwnerOriQ(first(), oriQ); }
void DEMTracker::AddAcc(float3 acc, size_t offset) { sys->AddOwnerNextStepAcc(GetOwnerID(offset), std::vector<float3>{acc}); }
void DEMTracker::AddAcc(const std::vector<float3>& acc) { assertOwnerSize(acc.size(), "AddAcc"); sys->AddOwnerNextStepAcc(first(), acc); }
void DEMTracker::AddAngAcc(float3 angAcc, size_t offset) { sys->AddOwnerNextStepAngAcc(GetOwnerID(offset), std::vector<float3>{angAcc}); }
void DEMTracker::AddAngAcc(const std::vector<float3>& angAcc) { assertOwnerSize(angAcc.size(), "AddAngAcc"); sys->AddOwnerNextStepAngAcc(first(), angAcc); }
void DEMTracker::SetFamily(unsigned int fam_num) { sys->SetOwnerFamily(first(), fam_num, (bodyID_t)count()); }
void DEMTracker::SetFamily(unsigned int fam_num, size_t offset) { sys->SetOwnerFamily(GetOwnerID(offset), fam_num); }
void DEMTracker::ChangeClumpSizes(const std::vector<bodyID_t>& IDs, const std::vector<float>& factors) {
    std::vector<bodyID_t> offsetted = IDs;
    for (bodyID_t& id : offsetted) id += first();
    sys->ChangeClumpSizes(offsetted, factors);
}
float DEMTracker::GetOwnerWildcardValue(const std::string& name, size_t offset) { return sys->GetOwnerWildcardValue(GetOwnerID(offset), name)[0]; }
std::vector<float> DEMTracker::GetOwnerWildcardValues(const std::string& name) { return sys->GetOwnerWildcardValue(first(), name, (bodyID_t)count()); }
float DEMTracker::GetGeometryWildcardValue(const std::string& name, size_t offset) { return sys->GetSphereWildcardValue((bodyID_t)offset, name, 1)[0]; }
std::vector<float> DEMTracker::GetGeometryWildcardValues(const std::string& name) { return sys->GetSphereWildcardValue(0, name, 1); }
void DEMTracker::SetOwnerWildcardValue(const std::string& name, float wc, size_t offset) { sys->SetOwnerWildcardValue(GetOwnerID(offset), name, wc); }
void DEMTracker::SetOwnerWildcardValues(const std::string& name, const std::vector<float>& wc) { sys->SetOwnerWildcardValue(first(), name, wc); }
void DEMTracker::SetGeometryWildcardValue(const std::string& name, float wc, size_t offset) { sys->SetSphereWildcardValue((bodyID_t)offset, name, std::vector<float>{wc}); }
void DEMTracker::SetGeometryWildcardValues(const std::string& name, const std::vector<float>& wc) { sys->SetSphereWildcardValue(0, name, wc); }

// ---- inspectors ----
DEMInspector::DEMInspector(DEMSolver* sim, const std::string& quantity) : sys(sim) {
    // (AuxClasses.cpp:88-164 of the reference: the first three look at every sphere, the others at the owners)
    if (quantity == "clump_max_z") kind = DEM_REDUCE_SPHERE_MAX_Z;
    else if (quantity == "clump_min_z") kind = DEM_REDUCE_SPHERE_MIN_Z;
    else if (quantity == "clump_max_absv") kind = DEM_REDUCE_SPHERE_MAX_ABSV;
    else if (quantity == "max_absv") kind = DEM_REDUCE_MAX_ABSV;  // every owner, not only clumps: see GetValue
    else if (quantity == "clump_kinetic_energy") kind = DEM_REDUCE_KINETIC_ENERGY;
    else if (quantity == "clump_mass") kind = DEM_REDUCE_TOTAL_MASS;
    else if (quantity == "clump_volume") kind = 100;  // KIND_CLUMP_VOLUME: summed on the host from the templates
    else if (quantity == "absv") kind = 102;          // KIND_ABSV_ALL: un-reduced, one value per owner (GetValues)
    else fail(quantity + " is not a known query type (available: clump_max_z, clump_min_z, clump_max_absv, max_absv, absv, "
              "clump_kinetic_energy, clump_mass, clump_volume).");
}
float* DEMInspector::GetValues() {
    if (kind != 102)
        fail("GetValues is for un-reduced quantities (absv); this inspector reduces its quantity to one number: use GetValue().");
    const auto v = sys->GetOwnerVelocity(0, (bodyID_t)sys->GetNumOwners());
    m_values.resize(v.size());
    for (size_t i = 0; i < v.size(); i++)
        m_values[i] = (float)std::sqrt((double)v[i].x * v[i].x + (double)v[i].y * v[i].y + (double)v[i].z * v[i].z);  // AuxClasses.cpp:54-61
    return m_values.data();
}
void DEMInspector::SetInspectionCode(const std::string&) {
    fail("SetInspectionCode: inspection code is C++ text the reference compiles into its query kernel at run time; this "
         "ahead-of-time compiled core offers the built-in quantities (optionally confined to a region) instead.");
}
float DEMInspector::GetValue() {
    if (kind == 102) return GetValues()[0];
    if (region) return (float)sys->ReduceInRegion(kind, *region);
    if (kind == 100) return (float)sys->ReduceInRegion(kind, ScalarExpression("1", {"X", "Y", "Z"}));
    double value = sys->Reduce(kind);
    if (kind == DEM_REDUCE_MAX_ABSV) {
        // "max_absv" looks at EVERYTHING (AuxClasses.cpp:143-149 of the reference): the device reduction covers the clumps,
        // the few external objects and meshes behind them are read back
        const size_t nC = sys->GetNumClumps(), nO = sys->GetNumOwners();
        if (nO > nC)
            for (const float3& v : sys->GetOwnerVelocity((bodyID_t)nC, (bodyID_t)(nO - nC)))
                value = std::max(value, std::sqrt((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z));
    }
    return (float)value;
}

// ---- writers (dT.cpp:1254-1617): CSV only ----
namespace {
// the per-owner columns both writers share, in the reference's order (dT.cpp:1266-1301 / :1503-1530)
struct OwnerColumns {
    std::vector<float> q, v, w, acc, angacc;
    std::vector<uint8_t> fam;
};
void owner_header(std::ostream& f, unsigned int content) {
    if (content & ABSV) f << ",absv";
    if (content & VEL) f << ",v_x,v_y,v_z";
    if (content & ANG_VEL) f << ",w_x,w_y,w_z";
    if (content & ABS_ACC) f << ",abs_acc";
    if (content & ACC) f << ",a_x,a_y,a_z";
    if (content & ANG_ACC) f << ",alpha_x,alpha_y,alpha_z";
    if (content & FAMILY) f << ",family";
}
void owner_row(std::ostream& f, unsigned int content, const OwnerColumns& c, size_t i) {
    const float* v = &c.v[3 * i];
    const float* w = &c.w[3 * i];
    const float* a = &c.acc[3 * i];
    const float* al = &c.angacc[3 * i];
    if (content & ABSV) f << "," << std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (content & VEL) f << "," << v[0] << "," << v[1] << "," << v[2];
    if (content & ANG_VEL) f << "," << w[0] << "," << w[1] << "," << w[2];
    if (content & ABS_ACC) f << "," << std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (content & ACC) f << "," << a[0] << "," << a[1] << "," << a[2];
    if (content & ANG_ACC) f << "," << al[0] << "," << al[1] << "," << al[2];
    if (content & FAMILY) f << "," << (unsigned)c.fam[i];
}
}  // namespace
void DEMSolver::WriteClumpFile(const std::filesystem::path& outfilename, unsigned int accuracy) const {
    assertInit("WriteClumpFile");
    warnIfBinary(m_out_format, "clump");
    const uint32_t n = (uint32_t)nOwnerClumps;
    std::vector<float> pos(3 * (size_t)n);
    OwnerColumns c;
    c.q.resize(4 * (size_t)n); c.v.resize(3 * (size_t)n); c.w.resize(3 * (size_t)n);
    c.acc.resize(3 * (size_t)n); c.angacc.resize(3 * (size_t)n); c.fam.resize(n);
    const bool want_acc = (m_out_content & (ABS_ACC | ACC | ANG_ACC)) != 0;
    check(dem_download_positions(ctx, 0, n, pos.data(), nullptr), "dem_download_positions");
    check(dem_download_owner_state(ctx, 0, n, nullptr, nullptr, nullptr, nullptr, c.q.data(), c.v.data(), c.w.data(),
                                   want_acc ? c.acc.data() : nullptr, want_acc ? c.angacc.data() : nullptr, c.fam.data()),
          "dem_download_owner_state");
    std::ofstream f(outfilename);
    f << std::setprecision(accuracy);
    // position, orientation and type always; the rest as SetOutputContent says (dT.cpp:1501-1531 of the reference)
    f << "X,Y,Z,Qw,Qx,Qy,Qz,clump_type";
    owner_header(f, m_out_content);
    f << "\n";
    for (uint32_t i = 0; i < n; i++) {
        if (m_no_output_families.count(c.fam[i])) continue;  // DisableFamilyOutput
        f << pos[3 * i] << "," << pos[3 * i + 1] << "," << pos[3 * i + 2];
        f << "," << c.q[4 * i] << "," << c.q[4 * i + 1] << "," << c.q[4 * i + 2] << "," << c.q[4 * i + 3];
        f << "," << m_templates[m_owner_type_mark[i]]->m_name;
        owner_row(f, m_out_content, c, i);
        f << "\n";
    }
}
void DEMSolver::WriteSphereFile(const std::filesystem::path& outfilename) const {
    assertInit("WriteSphereFile");
    warnIfBinary(m_out_format, "sphere");
    const uint32_t n = (uint32_t)nOwnerClumps;
    std::vector<float> pos(3 * (size_t)n);
    OwnerColumns c;
    c.q.resize(4 * (size_t)n); c.v.resize(3 * (size_t)n); c.w.resize(3 * (size_t)n);
    c.acc.resize(3 * (size_t)n); c.angacc.resize(3 * (size_t)n); c.fam.resize(n);
    const bool want_acc = (m_out_content & (ABS_ACC | ACC | ANG_ACC)) != 0;
    check(dem_download_positions(ctx, 0, n, pos.data(), nullptr), "dem_download_positions");
    check(dem_download_owner_state(ctx, 0, n, nullptr, nullptr, nullptr, nullptr, c.q.data(), c.v.data(), c.w.data(),
                                   want_acc ? c.acc.data() : nullptr, want_acc ? c.angacc.data() : nullptr, c.fam.data()),
          "dem_download_owner_state");
    std::ofstream f(outfilename);
    // one row per sphere: centre, radius, then its owner's columns (dT.cpp:1263-1303, 1340-1400 of the reference)
    f << "X,Y,Z,r";
    owner_header(f, m_out_content);
    f << "\n";
    for (uint32_t i = 0; i < n; i++) {
        if (m_no_output_families.count(c.fam[i])) continue;  // DisableFamilyOutput
        const auto& t = m_templates[m_owner_type_mark[i]];
        const float4 quat = make_float4(c.q[4 * i + 1], c.q[4 * i + 2], c.q[4 * i + 3], c.q[4 * i]);
        for (unsigned int k = 0; k < t->nComp; k++) {
            const float3 r = Rotate(t->relPos[k], quat);
            f << pos[3 * i] + r.x << "," << pos[3 * i + 1] + r.y << "," << pos[3 * i + 2] + r.z << "," << t->radii[k];
            owner_row(f, m_out_content, c, i);
            f << "\n";
        }
    }
}
void DEMSolver::WriteContactFile(const std::filesystem::path& outfilename, float force_thres) const {
    // the columns SetContactOutputContent selected, in the reference's order (writeContactsAsCsv, dT.cpp:1757-1847):
    // contact_type | A,B | geoA,geoB | f_x,f_y,f_z | X,Y,Z | n_x,n_y,n_z | torque_x,torque_y,torque_z | the wildcards
    assertInit("WriteContactFile");
    warnIfBinary(m_cnt_out_format, "contact pair");
    unsigned int content = m_cnt_out_content;
    if (no_recording_contact_forces) {  // without the force record there is nothing to threshold on or to report
        content &= ~(unsigned int)(FORCE | CNT_POINT | NORMAL | TORQUE);
        force_thres = -1.f;
    }
    const auto info = generateContactInfo(force_thres, content);
    std::ofstream f(outfilename);
    f << std::setprecision(9);
    f << "contact_type";
    if (content & OWNER) f << ",A,B";
    if (content & GEO_ID) f << ",geoA,geoB";
    if (content & FORCE) f << ",f_x,f_y,f_z";
    if (content & CNT_POINT) f << ",X,Y,Z";
    if (content & NORMAL) f << ",n_x,n_y,n_z";
    if (content & TORQUE) f << ",torque_x,torque_y,torque_z";
    if (content & CNT_WILDCARD)
        for (const std::string& name : info->GetWildcardNames()) f << "," << name;
    f << "\n";
    auto put3 = [&](const float3& v) { f << "," << v.x << "," << v.y << "," << v.z; };
    for (size_t i = 0; i < info->Size(); i++) {
        f << info->GetContactType()[i];
        if (content & OWNER) f << "," << info->GetAOwner()[i] << "," << info->GetBOwner()[i];
        if (content & GEO_ID) f << "," << info->GetAGeo()[i] << "," << info->GetBGeo()[i];
        if (content & FORCE) put3(info->GetForce()[i]);
        if (content & CNT_POINT) put3(info->GetPoint()[i]);
        if (content & NORMAL) put3(info->GetNormal()[i]);
        if (content & TORQUE) put3(info->GetTorque()[i]);
        if (content & CNT_WILDCARD)
            for (const std::string& name : info->GetWildcardNames()) f << "," << info->GetWildcard(name)[i];
        f << "\n";
    }
}

static std::vector<std::vector<std::string>> read_csv(const std::string& fn, std::vector<std::string>& header) {
    std::ifstream f(fn);
    if (!f) fail("File " + fn + " cannot be opened.");
    // files written on Windows end their lines with \r\n (data/clumps/ContactChain_initial.csv does); cells are trimmed
    auto trimmed = [](std::string t) {
        while (!t.empty() && isspace((unsigned char)t.back())) t.pop_back();
        size_t b = 0;
        while (b < t.size() && isspace((unsigned char)t[b])) b++;
        return t.substr(b);
    };
    std::string line;
    std::getline(f, line);
    std::stringstream hs(trimmed(line));
    std::string c;
    while (std::getline(hs, c, ',')) header.push_back(trimmed(c));
    std::vector<std::vector<std::string>> rows;
    while (std::getline(f, line)) {
        line = trimmed(line);
        if (line.empty()) continue;
        std::stringstream ss(line);
        std::vector<std::string> r;
        while (std::getline(ss, c, ',')) r.push_back(trimmed(c));
        rows.push_back(r);
    }
    return rows;
}
std::unordered_map<std::string, std::vector<float3>> DEMSolver::ReadClumpXyzFromCsv(
    const std::string& infilename, const std::string& clump_header, const std::string& x_header,
    const std::string& y_header, const std::string& z_header) {
    std::vector<std::string> h;
    auto rows = read_csv(infilename, h);
    auto col = [&](const std::string& id) { return (size_t)(std::find(h.begin(), h.end(), id) - h.begin()); };
    const size_t ic = col(clump_header), ix = col(x_header), iy = col(y_header), iz = col(z_header);
    if (std::max(std::max(ic, ix), std::max(iy, iz)) >= h.size()) fail("ReadClumpXyzFromCsv: a requested column is missing in " + infilename);
    std::unordered_map<std::string, std::vector<float3>> out;
    for (const auto& r : rows)
        out[r[ic]].push_back(make_float3((float)atof(r[ix].c_str()), (float)atof(r[iy].c_str()), (float)atof(r[iz].c_str())));
    return out;
}
std::unordered_map<std::string, std::vector<float4>> DEMSolver::ReadClumpQuatFromCsv(
    const std::string& infilename, const std::string& clump_header, const std::string& qw_header,
    const std::string& qx_header, const std::string& qy_header, const std::string& qz_header) {
    std::vector<std::string> h;
    auto rows = read_csv(infilename, h);
    auto col = [&](const std::string& id) { return (size_t)(std::find(h.begin(), h.end(), id) - h.begin()); };
    const size_t ic = col(clump_header), iw = col(qw_header), ix = col(qx_header), iy = col(qy_header), iz = col(qz_header);
    if (std::max(std::max(ic, iw), std::max(ix, std::max(iy, iz))) >= h.size())
        fail("ReadClumpQuatFromCsv: a requested column is missing in " + infilename);
    std::unordered_map<std::string, std::vector<float4>> out;
    for (const auto& r : rows)
        out[r[ic]].push_back(make_float4((float)atof(r[ix].c_str()), (float)atof(r[iy].c_str()), (float)atof(r[iz].c_str()),
                                         (float)atof(r[iw].c_str())));
    return out;
}

// ---- deforming meshes ----
std::shared_ptr<DEMMeshConnected> DEMSolver::GetCachedMesh(bodyID_t ownerID) {
    for (auto& m : m_cached_meshes)
        if (m->owner == ownerID) return m;
    fail("Owner " + std::to_string(ownerID) + " is not a mesh.");
}
void DEMSolver::SetTriNodeRelPos(size_t owner, size_t triID, const std::vector<float3>& new_nodes) {
    assertInit("SetTriNodeRelPos");
    auto me = GetCachedMesh((bodyID_t)owner);
    if (new_nodes.size() != me->m_vertices.size())
        fail("SetTriNodeRelPos: the mesh has " + std::to_string(me->m_vertices.size()) + " nodes, but " +
             std::to_string(new_nodes.size()) + " new node positions were given.");
    if (triID != me->tri_first) fail("SetTriNodeRelPos: triID must be the id of the mesh's first facet.");
    me->m_vertices = new_nodes;
    std::vector<float> n1, n2, n3;
    n1.reserve(3 * me->nTri); n2.reserve(3 * me->nTri); n3.reserve(3 * me->nTri);
    for (size_t t = 0; t < me->nTri; t++) {
        const int3 fc = me->m_face_v_indices[t];
        const float3 a = new_nodes[fc.x], b = new_nodes[fc.y], c = new_nodes[fc.z];
        n1.insert(n1.end(), {a.x, a.y, a.z});
        n2.insert(n2.end(), {b.x, b.y, b.z});
        n3.insert(n3.end(), {c.x, c.y, c.z});
    }
    check(dem_update_triangle_nodes(ctx, (uint32_t)me->tri_first, (uint32_t)me->nTri, n1.data(), n2.data(), n3.data()),
          "dem_update_triangle_nodes");
}
void DEMSolver::UpdateTriNodeRelPos(size_t owner, size_t triID, const std::vector<float3>& updates) {
    assertInit("UpdateTriNodeRelPos");
    auto me = GetCachedMesh((bodyID_t)owner);
    if (updates.size() != me->m_vertices.size())
        fail("UpdateTriNodeRelPos: the mesh has " + std::to_string(me->m_vertices.size()) + " nodes, but " +
             std::to_string(updates.size()) + " increments were given.");
    std::vector<float3> moved = me->m_vertices;
    for (size_t i = 0; i < moved.size(); i++) moved[i] += updates[i];
    SetTriNodeRelPos(owner, triID, moved);
}
std::vector<float3> DEMSolver::GetMeshNodesGlobal(bodyID_t ownerID) {
    assertInit("GetMeshNodesGlobal");
    auto me = GetCachedMesh(ownerID);
    const float3 pos = GetOwnerPosition(ownerID)[0];
    const float4 q = GetOwnerOriQ(ownerID)[0];
    std::vector<float3> out(me->m_vertices.size());
    for (size_t i = 0; i < out.size(); i++) out[i] = Rotate(me->m_vertices[i], q) + pos;
    return out;
}
static std::shared_ptr<DEMMeshConnected> tracked_mesh(const std::shared_ptr<DEMInitializer>& obj, const char* what) {
    if (obj->obj_type != OWNER_TYPE::MESH) fail(std::string(what) + " is only callable when this tracker is tracking a mesh.");
    return std::static_pointer_cast<DEMMeshConnected>(obj);
}
void DEMTracker::UpdateMesh(const std::vector<float3>& new_nodes) {
    auto me = tracked_mesh(obj, "UpdateMesh");
    sys->SetTriNodeRelPos(me->owner, me->tri_first, new_nodes);
}
void DEMTracker::UpdateMeshByIncrement(const std::vector<float3>& deformation) {
    auto me = tracked_mesh(obj, "UpdateMeshByIncrement");
    sys->UpdateTriNodeRelPos(me->owner, me->tri_first, deformation);
}
std::vector<float3> DEMTracker::GetMeshNodesGlobal() { return sys->GetMeshNodesGlobal(tracked_mesh(obj, "GetMeshNodesGlobal")->owner); }
std::shared_ptr<DEMMeshConnected> DEMTracker::GetMesh() { return tracked_mesh(obj, "GetMesh"); }

// ---- contact queries ----
namespace {
struct ContactRows {
    std::vector<uint32_t> a, b;
    std::vector<uint8_t> t;
    std::vector<float> force, point;
};
}  // namespace
static ContactRows download_rows(DemCtx* ctx, bool with_record) {
    ContactRows r;
    uint64_t n = 0;
    if (dem_download_contacts(ctx, 0, &n, nullptr, nullptr, nullptr, nullptr, nullptr) != DEM_OK) fail(dem_last_error(ctx));
    r.a.resize(n); r.b.resize(n); r.t.resize(n);
    if (with_record) { r.force.resize(3 * n); r.point.resize(3 * n); }
    if (n && dem_download_contact_records(ctx, n, &n, r.a.data(), r.b.data(), r.t.data(), nullptr,
                                          with_record ? r.force.data() : nullptr, with_record ? r.point.data() : nullptr) != DEM_OK)
        fail(dem_last_error(ctx));
    return r;
}
bodyID_t DEMSolver::geoOwner(uint32_t geo, uint8_t type, bool sideB) const {
    if (!sideB || type == DEM_CNT_SPHERE_SPHERE) return m_sphere_owner.at(geo);
    if (type == DEM_CNT_SPHERE_MESH) return m_tri_owner.at(geo);
    return m_anal_owner.at(geo);
}
std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::contactOwnerPairs(bool clumps_only, const std::set<family_t>* fams) const {
    assertInit("GetContacts");
    ContactRows r = download_rows(ctx, false);
    if (!m_persistent.empty()) {  // marked pairs the list has dropped since stay reported
        std::set<std::tuple<uint32_t, uint32_t, uint8_t>> listed;
        for (size_t i = 0; i < r.a.size(); i++) listed.emplace(r.a[i], r.b[i], r.t[i]);
        for (const auto& p : m_persistent)
            if (!listed.count(std::make_tuple(p.geoA, p.geoB, p.type))) { r.a.push_back(p.geoA); r.b.push_back(p.geoB); r.t.push_back(p.type); }
    }
    std::vector<uint8_t> fam;
    if (fams) {
        fam.resize(nOwnerBodies);
        check(dem_download_owner_state(ctx, 0, (uint32_t)nOwnerBodies, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                       nullptr, nullptr, nullptr, fam.data()), "dem_download_owner_state");
    }
    std::vector<std::pair<bodyID_t, bodyID_t>> out;
    for (size_t i = 0; i < r.a.size(); i++) {
        if (clumps_only && r.t[i] != DEM_CNT_SPHERE_SPHERE) continue;
        const bodyID_t oa = geoOwner(r.a[i], r.t[i], false), ob = geoOwner(r.b[i], r.t[i], true);
        if (fams && (!fams->count(fam[oa]) || !fams->count(fam[ob]))) continue;
        out.emplace_back(oa, ob);
    }
    std::stable_sort(out.begin(), out.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    return out;
}
std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::GetContacts() const { return contactOwnerPairs(false, nullptr); }
std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::GetContacts(const std::set<family_t>& f) const { return contactOwnerPairs(false, &f); }
std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::GetClumpContacts() const { return contactOwnerPairs(true, nullptr); }
std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::GetClumpContacts(const std::set<family_t>& f) const { return contactOwnerPairs(true, &f); }

size_t DEMSolver::GetOwnerContactForces(const std::vector<bodyID_t>& ownerIDs, std::vector<float3>& points,
                                        std::vector<float3>& forces) const {
    assertInit("GetOwnerContactForces");
    if (no_recording_contact_forces)
        fail("GetOwnerContactForces needs the per-contact force record; do not call SetNoForceRecord() if you query contact forces.");
    const std::set<bodyID_t> want(ownerIDs.begin(), ownerIDs.end());
    const ContactRows r = download_rows(ctx, true);
    points.clear();
    forces.clear();
    for (size_t i = 0; i < r.a.size(); i++) {
        const float3 F = make_float3(r.force[3 * i], r.force[3 * i + 1], r.force[3 * i + 2]);
        if (length(F) < 1e-15f) continue;  // DEME_TINY_FLOAT: only contacts that produce force
        const bodyID_t oa = geoOwner(r.a[i], r.t[i], false), ob = geoOwner(r.b[i], r.t[i], true);
        const bool forA = want.count(oa) != 0;
        if (!forA && !want.count(ob)) continue;
        points.push_back(make_float3(r.point[3 * i], r.point[3 * i + 1], r.point[3 * i + 2]));
        forces.push_back(forA ? F : F * -1.f);
    }
    return points.size();
}
size_t DEMTracker::GetContactForces(std::vector<float3>& points, std::vector<float3>& forces, size_t offset) {
    return sys->GetOwnerContactForces({GetOwnerID(offset)}, points, forces);
}
size_t DEMTracker::GetContactForcesForAll(std::vector<float3>& points, std::vector<float3>& forces) {
    std::vector<bodyID_t> ids;
    const size_t n = (obj->obj_type == OWNER_TYPE::CLUMP) ? std::static_pointer_cast<DEMClumpBatch>(obj)->GetNumClumps() : 1;
    for (size_t i = 0; i < n; i++) ids.push_back(GetOwnerID(i));
    return sys->GetOwnerContactForces(ids, points, forces);
}

std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::ReadContactPairsFromCsv(const std::string& infilename,
                                                                            const std::string& cntType,
                                                                            const std::string& cntColName,
                                                                            const std::string& first_name,
                                                                            const std::string& second_name) {
    std::vector<std::string> h;
    auto rows = read_csv(infilename, h);
    auto col = [&](const std::string& id) { return (size_t)(std::find(h.begin(), h.end(), id) - h.begin()); };
    const size_t it = col(cntColName), ia = col(first_name), ib = col(second_name);
    if (std::max(it, std::max(ia, ib)) >= h.size()) fail("ReadContactPairsFromCsv: a requested column is missing in " + infilename);
    std::vector<std::pair<bodyID_t, bodyID_t>> pairs;
    for (const auto& r : rows)
        if (r[it] == cntType) pairs.emplace_back((bodyID_t)std::stoul(r[ia]), (bodyID_t)std::stoul(r[ib]));
    return pairs;
}
std::unordered_map<std::string, std::vector<float>> DEMSolver::ReadContactWildcardsFromCsv(const std::string& infilename,
                                                                                         const std::string& cntType,
                                                                                         const std::string& cntColName) {
    // every column that is not one of the standard contact-file columns is a wildcard (Structs.h:79-87 of the reference)
    static const std::set<std::string> known = {"A", "B", "compA", "compB", "geoA", "geoB", "nameA", "nameB", "contact_type",
                                                "f_x", "f_y", "f_z", "torque_x", "torque_y", "torque_z", "n_x", "n_y", "n_z",
                                                "X", "Y", "Z", "SS", "SA", "SM"};
    std::vector<std::string> h;
    auto rows = read_csv(infilename, h);
    const size_t it = (size_t)(std::find(h.begin(), h.end(), cntColName) - h.begin());
    if (it >= h.size()) fail("ReadContactWildcardsFromCsv: column " + cntColName + " is missing in " + infilename);
    std::unordered_map<std::string, std::vector<float>> out;
    for (size_t c = 0; c < h.size(); c++) {
        if (known.count(h[c])) continue;
        auto& v = out[h[c]];
        for (const auto& r : rows)
            if (r[it] == cntType) v.push_back((float)atof(r[c].c_str()));
    }
    return out;
}


// ---------------------------------------------------------------------------------------------------------------
// Contact read-outs with the fields of SetContactOutputContent, contact wildcards, persistent contacts
// ---------------------------------------------------------------------------------------------------------------
namespace {
// one row per listed contact: geometry ids, type, force on A + contact point (world frame), the four history words
struct FullRows {
    std::vector<uint32_t> a, b;
    std::vector<uint8_t> t;
    std::vector<float> force, point, wc;
    size_t size() const { return a.size(); }
};
FullRows download_full_rows(DemCtx* ctx, bool with_record) {
    FullRows r;
    uint64_t n = 0;
    if (dem_download_contacts(ctx, 0, &n, nullptr, nullptr, nullptr, nullptr, nullptr) != DEM_OK) fail(dem_last_error(ctx));
    r.a.resize(n); r.b.resize(n); r.t.resize(n); r.wc.resize(4 * n);
    r.force.assign(3 * n, 0.f); r.point.assign(3 * n, 0.f);
    if (!n) return r;
    const int rc = with_record ? dem_download_contact_records(ctx, n, &n, r.a.data(), r.b.data(), r.t.data(), r.wc.data(),
                                                              r.force.data(), r.point.data())
                               : dem_download_contacts(ctx, n, &n, r.a.data(), r.b.data(), r.t.data(), r.wc.data(), nullptr);
    if (rc != DEM_OK) fail(dem_last_error(ctx));
    return r;
}
const char* contact_type_name(uint8_t t) {
    return t == DEM_CNT_SPHERE_SPHERE ? "SS" : (t == DEM_CNT_SPHERE_MESH ? "SM" : "SA");
}
const char* const kWildcardNames[4] = {"delta_tan_x", "delta_tan_y", "delta_tan_z", "delta_time"};
}  // namespace

ContactInfoContainer::ContactInfoContainer(unsigned int cnt_out_content, const std::vector<std::string>& wildcard_names)
    : m_content(cnt_out_content), m_wc_names(wildcard_names), m_wc(wildcard_names.size()) {}
std::vector<float>& ContactInfoContainer::GetWildcard(const std::string& name) {
    if (m_content & CNT_WILDCARD)
        for (size_t k = 0; k < m_wc_names.size(); k++)
            if (m_wc_names[k] == name) return m_wc[k];
    throw std::runtime_error("ContactInfoContainer does not have field: '" + name +
                             "', you may need to turn on the output of this field by correctly calling "
                             "SetContactOutputContent before Initialize().");
}
void ContactInfoContainer::ResizeAll(size_t n) {
    m_type.resize(n);
    m_family[0].resize(n); m_family[1].resize(n);
    if (m_content & CNT_POINT) m_point.resize(n);
    if (m_content & FORCE) m_force.resize(n);
    if (m_content & TORQUE) m_torque.resize(n);
    if (m_content & NORMAL) m_normal.resize(n);
    if (m_content & OWNER) { m_owner[0].resize(n); m_owner[1].resize(n); }
    if (m_content & GEO_ID) { m_geo[0].resize(n); m_geo[1].resize(n); }
    if (m_content & CNT_WILDCARD)
        for (auto& w : m_wc) w.resize(n);
}

// marked pairs the broad phase no longer lists come back as rows without force (see the header)
static void append_persistent(FullRows& r, const std::vector<std::tuple<uint32_t, uint32_t, uint8_t>>& marked) {
    if (marked.empty()) return;
    std::set<std::tuple<uint32_t, uint32_t, uint8_t>> listed;
    for (size_t i = 0; i < r.size(); i++) listed.emplace(r.a[i], r.b[i], r.t[i]);
    for (const auto& m : marked) {
        if (listed.count(m)) continue;
        r.a.push_back(std::get<0>(m)); r.b.push_back(std::get<1>(m)); r.t.push_back(std::get<2>(m));
        r.force.insert(r.force.end(), 3, 0.f);
        r.point.insert(r.point.end(), 3, 0.f);
        r.wc.insert(r.wc.end(), 4, 0.f);
    }
}
std::vector<std::tuple<uint32_t, uint32_t, uint8_t>> DEMSolver::persistentKeys() const {
    std::vector<std::tuple<uint32_t, uint32_t, uint8_t>> keys;
    for (const auto& p : m_persistent) keys.emplace_back(p.geoA, p.geoB, p.type);
    return keys;
}

std::shared_ptr<ContactInfoContainer> DEMSolver::GetContactDetailedInfo(float force_thres) const {
    assertInit("GetContactDetailedInfo");
    return generateContactInfo(force_thres, m_cnt_out_content);
}
std::shared_ptr<ContactInfoContainer> DEMSolver::generateContactInfo(float force_thres, unsigned int content) const {
    const bool has_record = !no_recording_contact_forces;
    if (!has_record && (content & (FORCE | CNT_POINT | NORMAL | TORQUE)))
        fail("GetContactDetailedInfo: force, point, normal and torque come from the per-contact force record; do not call "
             "SetNoForceRecord() if you query them.");
    FullRows r = download_full_rows(ctx, has_record);
    append_persistent(r, persistentKeys());
    std::vector<std::string> names;
    if (m_force_model == FORCE_MODEL::HERTZIAN) names.assign(kWildcardNames, kWildcardNames + 4);
    auto info = std::make_shared<ContactInfoContainer>(content, names);
    info->ResizeAll(r.size());
    const uint32_t nO = (uint32_t)nOwnerBodies;
    std::vector<uint8_t> fam(nO);
    std::vector<float> q;
    std::vector<double> pos;
    check(dem_download_owner_state(ctx, 0, nO, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   fam.data()), "dem_download_owner_state");
    const bool want_normal = (content & NORMAL) != 0;
    if (want_normal) {
        q.resize(4 * (size_t)nO);
        pos.resize(3 * (size_t)nO);
        check(dem_download_owner_state(ctx, 0, nO, nullptr, nullptr, nullptr, nullptr, q.data(), nullptr, nullptr, nullptr,
                                       nullptr, nullptr), "dem_download_owner_state");
        check(dem_download_positions(ctx, 0, nO, nullptr, pos.data()), "dem_download_positions");
    }
    // the rolling-resistance couple of a contact is applied to the owners but not kept per contact
    bool warned = false;
    size_t useful = 0;
    for (size_t i = 0; i < r.size(); i++) {
        const float3 F = make_float3(r.force[3 * i], r.force[3 * i + 1], r.force[3 * i + 2]);
        if (length(F) < force_thres) continue;  // (dT.cpp:1647-1651 of the reference)
        const bodyID_t oa = geoOwner(r.a[i], r.t[i], false), ob = geoOwner(r.b[i], r.t[i], true);
        info->GetContactType()[useful] = contact_type_name(r.t[i]);
        info->GetAOwnerFamily()[useful] = fam[oa];
        info->GetBOwnerFamily()[useful] = fam[ob];
        if (content & OWNER) { info->GetAOwner()[useful] = oa; info->GetBOwner()[useful] = ob; }
        if (content & GEO_ID) { info->GetAGeo()[useful] = r.a[i]; info->GetBGeo()[useful] = r.b[i]; }
        if (content & FORCE) info->GetForce()[useful] = F;
        const float3 P = make_float3(r.point[3 * i], r.point[3 * i + 1], r.point[3 * i + 2]);
        if (content & CNT_POINT) info->GetPoint()[useful] = P;
        if (want_normal) {
            // outward normal of body A: from the centre of sphere A to the contact point (dT.cpp:1714-1727)
            const auto& tp = m_templates[m_owner_type_mark[oa]];
            const float4 qa = make_float4(q[4 * oa + 1], q[4 * oa + 2], q[4 * oa + 3], q[4 * oa]);
            const float3 off = Rotate(tp->relPos[r.a[i] - m_owner_first_sphere[oa]], qa);
            const float3 c = make_float3((float)pos[3 * oa] + off.x, (float)pos[3 * oa + 1] + off.y, (float)pos[3 * oa + 2] + off.z);
            const float3 d = P - c;
            const float len = length(d);
            info->GetNormal()[useful] = len > 0.f ? d * (1.f / len) : make_float3(0, 0, 0);
        }
        if (content & TORQUE) {
            if (!warned && m_any_rolling_resistance && verbosity >= WARNING) {
                std::cerr << "WARNING! The torque field of GetContactDetailedInfo is the couple a contact adds beyond the "
                             "moment of its force (rolling resistance). This core applies it to the owners without keeping "
                             "it per contact, so the field reads zero." << std::endl;
                warned = true;
            }
            info->GetTorque()[useful] = make_float3(0, 0, 0);
        }
        if (content & CNT_WILDCARD)
            for (size_t k = 0; k < names.size(); k++) info->GetWildcard(names[k])[useful] = r.wc[4 * i + k];
        useful++;
    }
    info->ResizeAll(useful);
    return info;
}

std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::GetContacts(std::vector<std::pair<family_t, family_t>>& family_pair) const {
    const auto pairs = contactOwnerPairs(false, nullptr);
    const auto fam = GetOwnerFamily(0, (bodyID_t)nOwnerBodies);
    family_pair.clear();
    for (const auto& p : pairs) family_pair.emplace_back((family_t)fam[p.first], (family_t)fam[p.second]);
    return pairs;
}
std::vector<std::pair<bodyID_t, bodyID_t>> DEMSolver::GetClumpContacts(std::vector<std::pair<family_t, family_t>>& family_pair) const {
    const auto pairs = contactOwnerPairs(true, nullptr);
    const auto fam = GetOwnerFamily(0, (bodyID_t)nOwnerBodies);
    family_pair.clear();
    for (const auto& p : pairs) family_pair.emplace_back((family_t)fam[p.first], (family_t)fam[p.second]);
    return pairs;
}

size_t DEMSolver::GetOwnerContactForces(const std::vector<bodyID_t>& ownerIDs, std::vector<float3>& points,
                                        std::vector<float3>& forces, std::vector<float3>& torques, bool torque_in_local) const {
    // (dT.cpp:2793-2853 of the reference: the torque is the contact's "torque-only force" -- the rolling-resistance couple --
    // turned into a moment about the owner's centre; the moment of the force itself is NOT part of it)
    (void)torque_in_local;  // a zero vector reads the same in both frames
    const size_t n = GetOwnerContactForces(ownerIDs, points, forces);
    if (m_any_rolling_resistance && verbosity >= WARNING)
        std::cerr << "WARNING! GetOwnerContactForces: the rolling-resistance couple is applied to the owners without being "
                     "kept per contact, so the torques read zero." << std::endl;
    torques.assign(n, make_float3(0, 0, 0));
    return n;
}
size_t DEMTracker::GetContactForcesAndGlobalTorque(std::vector<float3>& points, std::vector<float3>& forces,
                                                   std::vector<float3>& torques, size_t offset) {
    return sys->GetOwnerContactForces({GetOwnerID(offset)}, points, forces, torques, false);
}
size_t DEMTracker::GetContactForcesAndGlobalTorqueForAll(std::vector<float3>& points, std::vector<float3>& forces,
                                                         std::vector<float3>& torques) {
    return sys->GetOwnerContactForces(GetOwnerIDs(), points, forces, torques, false);
}
size_t DEMTracker::GetContactForcesAndLocalTorque(std::vector<float3>& points, std::vector<float3>& forces,
                                                  std::vector<float3>& torques, size_t offset) {
    return sys->GetOwnerContactForces({GetOwnerID(offset)}, points, forces, torques, true);
}
size_t DEMTracker::GetContactForcesAndLocalTorqueForAll(std::vector<float3>& points, std::vector<float3>& forces,
                                                        std::vector<float3>& torques) {
    return sys->GetOwnerContactForces(GetOwnerIDs(), points, forces, torques, true);
}

// ---- contact wildcards (the history words) ----
void DEMSolver::noSuchWildcard(const char* what, const std::string& name) const {
    fail(std::string("No ") + what + " wildcard in the force model is named " + name + ".\nThe built-in force models declare "
         "the contact wildcards delta_tan_x, delta_tan_y, delta_tan_z, delta_time (frictional model) and no owner or geometry "
         "wildcards; others belong to custom force models, which need run-time compilation that this core does not have.");
}
void DEMSolver::setContactWildcard(int mode, unsigned int N1, unsigned int N2, const std::string& name, float val) {
    int word = -1;
    if (m_force_model == FORCE_MODEL::HERTZIAN)
        for (int k = 0; k < 4; k++)
            if (name == kWildcardNames[k]) word = k;
    if (word < 0) noSuchWildcard("contact", name);
    // (dT.cpp:2855-2884 of the reference rewrites the word of every contact in its array whose owners' families match;
    // here the list is read back, edited and handed to the next rebuild as its history source, like a restart)
    FullRows r = download_full_rows(ctx, false);
    if (!r.size()) return;
    const auto fam = GetOwnerFamily(0, (bodyID_t)nOwnerBodies);
    size_t changed = 0;
    for (size_t i = 0; i < r.size(); i++) {
        const unsigned int fa = fam[geoOwner(r.a[i], r.t[i], false)], fb = fam[geoOwner(r.b[i], r.t[i], true)];
        bool hit = true;
        if (mode == 1) hit = (fa == N1 || fb == N1);
        else if (mode == 2) hit = (fa == N1 && fb == N1);
        else if (mode == 3) hit = (fa == N1 && fb == N2) || (fa == N2 && fb == N1);
        if (!hit) continue;
        r.wc[4 * i + word] = val;
        changed++;
    }
    if (!changed) return;
    check(dem_set_contacts(ctx, r.size(), r.a.data(), r.b.data(), r.t.data(), r.wc.data()), "dem_set_contacts");
    // the edited list is a history source only: rebuild now, so that contact queries made before the next step see a list
    check(dem_rebuild_contacts(ctx), "dem_rebuild_contacts");
}
void DEMSolver::SetContactWildcardValue(const std::string& name, float val) {
    assertInit("SetContactWildcardValue");
    setContactWildcard(0, 0, 0, name, val);
}
void DEMSolver::SetFamilyContactWildcardValueEither(unsigned int N, const std::string& name, float val) {
    assertInit("SetFamilyContactWildcardValueEither");
    setContactWildcard(1, N, 0, name, val);
}
void DEMSolver::SetFamilyContactWildcardValueBoth(unsigned int N, const std::string& name, float val) {
    assertInit("SetFamilyContactWildcardValueBoth");
    setContactWildcard(2, N, 0, name, val);
}
void DEMSolver::SetFamilyContactWildcardValue(unsigned int N1, unsigned int N2, const std::string& name, float val) {
    assertInit("SetFamilyContactWildcardValue");
    setContactWildcard(3, N1, N2, name, val);
}
void DEMSolver::SetContactWildcards(const std::set<std::string>& wildcards) { m_force_model_obj->SetPerContactWildcards(wildcards); }
void DEMSolver::SetOwnerWildcards(const std::set<std::string>& wildcards) { m_force_model_obj->SetPerOwnerWildcards(wildcards); }
void DEMSolver::SetGeometryWildcards(const std::set<std::string>& wildcards) { m_force_model_obj->SetPerGeometryWildcards(wildcards); }

// owner / geometry wildcards: no built-in model declares any, so every name is unknown (APIPublic.cpp:1042-1180)
void DEMSolver::SetOwnerWildcardValue(bodyID_t, const std::string& name, const std::vector<float>&) {
    assertInit("SetOwnerWildcardValue");
    noSuchWildcard("owner", name);
}
void DEMSolver::SetFamilyOwnerWildcardValue(unsigned int, const std::string& name, const std::vector<float>&) {
    assertInit("SetFamilyOwnerWildcardValue");
    noSuchWildcard("owner", name);
}
void DEMSolver::SetTriWildcardValue(bodyID_t, const std::string& name, const std::vector<float>&) {
    assertInit("SetTriWildcardValue");
    noSuchWildcard("geometry", name);
}
void DEMSolver::SetSphereWildcardValue(bodyID_t, const std::string& name, const std::vector<float>&) {
    assertInit("SetSphereWildcardValue");
    noSuchWildcard("geometry", name);
}
void DEMSolver::SetAnalWildcardValue(bodyID_t, const std::string& name, const std::vector<float>&) {
    assertInit("SetAnalWildcardValue");
    noSuchWildcard("geometry", name);
}
std::vector<float> DEMSolver::GetOwnerWildcardValue(bodyID_t, const std::string& name, bodyID_t) {
    assertInit("GetOwnerWildcardValue");
    noSuchWildcard("owner", name);
}
std::vector<float> DEMSolver::GetAllOwnerWildcardValue(const std::string& name) {
    assertInit("GetAllOwnerWildcardValue");
    noSuchWildcard("owner", name);
}
std::vector<float> DEMSolver::GetFamilyOwnerWildcardValue(unsigned int, const std::string& name) {
    assertInit("GetFamilyOwnerWildcardValue");
    noSuchWildcard("owner", name);
}
std::vector<float> DEMSolver::GetTriWildcardValue(bodyID_t, const std::string& name, size_t) {
    assertInit("GetTriWildcardValue");
    noSuchWildcard("geometry", name);
}
std::vector<float> DEMSolver::GetSphereWildcardValue(bodyID_t, const std::string& name, size_t) {
    assertInit("GetSphereWildcardValue");
    noSuchWildcard("geometry", name);
}
std::vector<float> DEMSolver::GetAnalWildcardValue(bodyID_t, const std::string& name, size_t) {
    assertInit("GetAnalWildcardValue");
    noSuchWildcard("geometry", name);
}

// ---- persistent contacts ----
void DEMSolver::markPersistent(int mode, unsigned int N1, unsigned int N2, bool mark) {
    if (m_force_model != FORCE_MODEL::HERTZIAN)
        fail("Persistent contacts need a force model with contact history (wildcards); the frictionless model has none.");
    // every pair that is in the list now, plus the marked ones the list has dropped since
    FullRows r = download_full_rows(ctx, false);
    append_persistent(r, persistentKeys());
    const auto fam = GetOwnerFamily(0, (bodyID_t)nOwnerBodies);
    for (size_t i = 0; i < r.size(); i++) {
        PersistentPair p;
        p.geoA = r.a[i]; p.geoB = r.b[i]; p.type = r.t[i];
        p.ownerA = geoOwner(r.a[i], r.t[i], false);
        p.ownerB = geoOwner(r.b[i], r.t[i], true);
        const unsigned int fa = fam[p.ownerA], fb = fam[p.ownerB];
        bool hit = true;
        if (mode == 1) hit = (fa == N1 || fb == N1);
        else if (mode == 2) hit = (fa == N1 && fb == N1);
        else if (mode == 3) hit = (fa == N1 && fb == N2) || (fa == N2 && fb == N1);
        if (!hit) continue;
        if (mark) m_persistent.insert(p);
        else m_persistent.erase(p);
    }
}
void DEMSolver::MarkFamilyPersistentContactEither(unsigned int N) { assertInit("MarkFamilyPersistentContactEither"); markPersistent(1, N, 0, true); }
void DEMSolver::MarkFamilyPersistentContactBoth(unsigned int N) { assertInit("MarkFamilyPersistentContactBoth"); markPersistent(2, N, 0, true); }
void DEMSolver::MarkFamilyPersistentContact(unsigned int N1, unsigned int N2) { assertInit("MarkFamilyPersistentContact"); markPersistent(3, N1, N2, true); }
void DEMSolver::MarkPersistentContact() { assertInit("MarkPersistentContact"); markPersistent(0, 0, 0, true); }
void DEMSolver::RemoveFamilyPersistentContactEither(unsigned int N) { assertInit("RemoveFamilyPersistentContactEither"); markPersistent(1, N, 0, false); }
void DEMSolver::RemoveFamilyPersistentContactBoth(unsigned int N) { assertInit("RemoveFamilyPersistentContactBoth"); markPersistent(2, N, 0, false); }
void DEMSolver::RemoveFamilyPersistentContact(unsigned int N1, unsigned int N2) { assertInit("RemoveFamilyPersistentContact"); markPersistent(3, N1, N2, false); }
void DEMSolver::RemovePersistentContact() { assertInit("RemovePersistentContact"); markPersistent(0, 0, 0, false); }

// ---- code-string "corrections" of the integrator (API.h:806-838 of the reference) ----
namespace {
[[noreturn]] void no_corrections(const char* who) {
    fail(std::string(who) + ": a correction is C++ code the reference compiles into its integration kernel at run time; this "
         "ahead-of-time compiled core has no run-time compilation. Prescribe the motion (SetFamilyPrescribedLinVel / "
         "...AngVel / ...Position, expressions of t are supported) or add an acceleration (AddFamilyPrescribedAcc) instead.");
}
}  // namespace
void DEMSolver::CorrectFamilyLinVel(unsigned int, const std::string&, const std::string&, const std::string&, const std::string&) {
    no_corrections("CorrectFamilyLinVel");
}
void DEMSolver::CorrectFamilyAngVel(unsigned int, const std::string&, const std::string&, const std::string&, const std::string&) {
    no_corrections("CorrectFamilyAngVel");
}
void DEMSolver::CorrectFamilyPosition(unsigned int, const std::string&, const std::string&, const std::string&, const std::string&) {
    no_corrections("CorrectFamilyPosition");
}
void DEMSolver::CorrectFamilyQuaternion(unsigned int, const std::string&) { no_corrections("CorrectFamilyQuaternion"); }

// ---- housekeeping ----
size_t DEMSolver::GetHostMemUsageDynamic() const {
    size_t b = m_owner_mass.size() * sizeof(float) + m_owner_moi.size() * sizeof(float3) +
               (m_owner_type_mark.size() + m_sphere_owner.size() + m_tri_owner.size() + m_anal_owner.size() +
                m_owner_first_sphere.size()) * sizeof(unsigned int) +
               m_family_masks.size();
    for (const auto& m : m_cached_meshes) b += m->m_vertices.size() * sizeof(float3) + m->m_face_v_indices.size() * sizeof(int3);
    return b;
}
void DEMSolver::UpdateSimParams() {
    // (APIPublic.cpp:2313-2326 of the reference: re-derive the margin / bin policy and push the parameters again)
    assertInit("UpdateSimParams");
    DemSimParams sp;
    memcpy(&sp, m_sp_blob.data(), sizeof(sp));
    const float g[3] = {G.x, G.y, G.z};
    for (int k = 0; k < 3; k++) sp.G[k] = g[k];
    sp.h = (float)m_ts_size;
    sp.cd_update_freq = (uint32_t)m_cd_update_freq;
    if (m_adaptive_update_freq) {  // keep what the tuner has settled on
        DemStats st;
        check(dem_get_stats(ctx, &st), "dem_get_stats");
        if (st.cd_update_freq >= 1) sp.cd_update_freq = st.cd_update_freq;
    }
    sp.beta = m_expand_factor;
    sp.approxMaxVel = m_approx_max_vel;
    sp.expSafetyMulti = m_expand_safety_multi;
    sp.expSafetyAdder = m_expand_base_vel;
    sp.errOutVel = threshold_error_out_vel;
    check(dem_set_params(ctx, &sp), "dem_set_params");
    memcpy(m_sp_blob.data(), &sp, sizeof(sp));
    check(dem_set_option(ctx, "update_freq_max", (double)m_max_update_freq), "SetCDMaxUpdateFreq");
    check(dem_set_option(ctx, "adaptive_update_freq", m_adaptive_update_freq ? 1.0 : 0.0), "UseAdaptiveUpdateFreq");
}
void DEMSolver::SetAdaptiveTimeStepType(const std::string& type) {
    if (verbosity >= WARNING)
        std::cerr << "WARNING! SetAdaptiveTimeStepType is currently not implemented and has no effect, time step size is still fixed."
                  << std::endl;
    const std::string u = upper(type);
    if (u != "NONE" && u != "MAX_VEL" && u != "INT_DIFF")
        fail("Adaptive time step type " + type + " is unknown. Please select another via SetAdaptiveTimeStepType.");
}

// ---- inspectors confined to a region ----
namespace {
constexpr int KIND_CLUMP_VOLUME = 100;  // a facade-side quantity (no device reduction behind it)
}  // namespace
DEMInspector::DEMInspector(DEMSolver* sim, const std::string& quantity, const std::string& region_code)
    : DEMInspector(sim, quantity) {
    bool blank = true;
    for (char c : region_code) blank = blank && isspace((unsigned char)c);
    if (blank) return;
    region = std::make_shared<ScalarExpression>(region_code, std::vector<std::string>{"X", "Y", "Z"});
    if (region->IsConstant())
        fail("One of your insepctors is set to query a specific region, but the region condition \"" + region_code +
             "\" does not depend on X, Y or Z. It should be a condition on the position, e.g. \"return (X > 0) && (Z < 1);\".");
}
std::shared_ptr<DEMInspector> DEMSolver::CreateInspector(const std::string& quantity, const std::string& region) {
    return std::make_shared<DEMInspector>(this, quantity, region);
}
double DEMSolver::ReduceInRegion(int kind, const ScalarExpression& region) const {
    assertInit("inspector");
    const bool clumps_only = (kind != DEM_REDUCE_MAX_ABSV);  // "max_absv" looks at every owner
    const uint32_t n = (uint32_t)(clumps_only ? nOwnerClumps : nOwnerBodies);
    if (!n) return 0.0;
    std::vector<double> pos(3 * (size_t)n);
    std::vector<float> q(4 * (size_t)n), v(3 * (size_t)n), w(3 * (size_t)n);
    check(dem_download_positions(ctx, 0, n, nullptr, pos.data()), "dem_download_positions");
    check(dem_download_owner_state(ctx, 0, n, nullptr, nullptr, nullptr, nullptr, q.data(), v.data(), w.data(), nullptr, nullptr,
                                   nullptr), "dem_download_owner_state");
    auto inside = [&](float X, float Y, float Z) {
        const double xyz[3] = {X, Y, Z};
        return region.Eval(xyz) != 0.0;
    };
    const bool per_sphere = (kind == DEM_REDUCE_SPHERE_MAX_Z || kind == DEM_REDUCE_SPHERE_MIN_Z || kind == DEM_REDUCE_SPHERE_MAX_ABSV);
    double acc = (kind == DEM_REDUCE_SPHERE_MIN_Z) ? DEME_HUGE_FLOAT : (kind == DEM_REDUCE_SPHERE_MAX_Z) ? -DEME_HUGE_FLOAT : 0.0;
    for (uint32_t i = 0; i < n; i++) {
        const float3 vi = make_float3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        if (!per_sphere) {
            // the owner's position decides (DEMOwnerQueryKernels.cu:42-48)
            if (!inside((float)pos[3 * i], (float)pos[3 * i + 1], (float)pos[3 * i + 2])) continue;
            if (kind == DEM_REDUCE_MAX_ABSV) {
                acc = std::max(acc, std::sqrt((double)vi.x * vi.x + (double)vi.y * vi.y + (double)vi.z * vi.z));
            } else if (kind == DEM_REDUCE_TOTAL_MASS) {
                acc += (float)m_owner_mass[i];
            } else if (kind == KIND_CLUMP_VOLUME) {
                acc += m_templates[m_owner_type_mark[i]]->volume;
            } else {  // kinetic energy (AuxClasses.cpp:62-75)
                const float3 I = m_owner_moi[i];
                double ke = 0.5 * (double)m_owner_mass[i] * ((double)vi.x * vi.x + (double)vi.y * vi.y + (double)vi.z * vi.z);
                ke += 0.5 * ((double)I.x * w[3 * i] * w[3 * i] + (double)I.y * w[3 * i + 1] * w[3 * i + 1] +
                             (double)I.z * w[3 * i + 2] * w[3 * i + 2]);
                acc += (float)ke;
            }
            continue;
        }
        // sphere quantities: the sphere's centre decides (DEMSphereQueryKernels.cu:41-46)
        const auto& tp = m_templates[m_owner_type_mark[i]];
        const float4 qi = make_float4(q[4 * i + 1], q[4 * i + 2], q[4 * i + 3], q[4 * i]);
        const float3 wl = make_float3(w[3 * i], w[3 * i + 1], w[3 * i + 2]);
        for (unsigned int k = 0; k < tp->nComp; k++) {
            const float3 off = Rotate(tp->relPos[k], qi);
            const float X = (float)(pos[3 * i] + off.x), Y = (float)(pos[3 * i + 1] + off.y), Z = (float)(pos[3 * i + 2] + off.z);
            if (!inside(X, Y, Z)) continue;
            if (kind == DEM_REDUCE_SPHERE_MAX_Z) acc = std::max(acc, (double)(Z + tp->radii[k]));
            else if (kind == DEM_REDUCE_SPHERE_MIN_Z) acc = std::min(acc, (double)(Z - tp->radii[k]));
            else acc = std::max(acc, (double)length(Rotate(cross(wl, tp->relPos[k]), qi) + vi));  // AuxClasses.cpp:27-50
        }
    }
    return acc;
}

}  // namespace deme
