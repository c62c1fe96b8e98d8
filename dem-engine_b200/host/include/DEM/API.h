// DEM/API.h -- host-side mirror of the reference's public interface (src/DEM/API.h:50-1263, AuxClasses.h:93-420,
// BdrsAndObjs.h:68-228, Structs.h:560-933 of projectchrono/DEM-Engine) for the hot path this repository rebuilds.
//
// Same namespace, class and method names, argument meaning and error behaviour (std::runtime_error thrown on the
// calling thread), so that an existing demo script that uses clumps, analytical boundaries, families with constant
// prescriptions, trackers and the built-in inspectors drives the B200 core unchanged.  Everything below the class
// surface is new: user input is flattened on the host and handed to the C ABI of include/dem_b200.h
// (libdemcore.so, hand-written sm_100a kernels); there is no jitify, no worker-thread pair and no CPU fallback.
// API members that need runtime compilation of user strings (custom force models, non-constant prescriptions,
// family-change conditions) throw with a clear message.
#pragma once

#include <algorithm>
#include <cstdint>
#include <filesystem>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include <vector_functions.h>
#include <vector_types.h>

#include "HostSideHelpers.hpp"

struct DemCtx;

#include "utils/Expression.hpp"

namespace deme {

typedef unsigned int bodyID_t;
typedef uint8_t family_t;
typedef uint8_t notStupidBool_t;  // the reference's array-friendly bool (src/DEM/VariableTypes.h:58)
constexpr double PI = 3.1415926535897932385;
#ifndef DEME_TINY_FLOAT
    #define DEME_TINY_FLOAT 1e-12
#endif
#ifndef DEME_HUGE_FLOAT
    #define DEME_HUGE_FLOAT 1e15
#endif

enum VERBOSITY { QUIET = 0, ERR = 10, WARNING = 20, INFO = 30, STEP_ANOMALY = 32, STEP_METRIC = 35, DEBUG = 40, STEP_DEBUG = 50 };
enum class TIME_INTEGRATOR { FORWARD_EULER, CENTERED_DIFFERENCE, EXTENDED_TAYLOR, CHUNG };
enum class OWNER_TYPE { CLUMP, ANALYTICAL, MESH };
enum class FORCE_MODEL { HERTZIAN, HERTZIAN_FRICTIONLESS, CUSTOM };
enum class OUTPUT_FORMAT { CSV, BINARY, CHPF };
enum class MESH_FORMAT { VTK, OBJ };
enum OUTPUT_CONTENT { XYZ = 0, QUAT = 1, ABSV = 2, VEL = 4, ANG_VEL = 8, ABS_ACC = 16, ACC = 32, ANG_ACC = 64, FAMILY = 128,
                      MAT = 256, OWNER_WILDCARD = 512, GEO_WILDCARD = 1024, EXP_FACTOR = 2048 };
enum CNT_OUTPUT_CONTENT { CNT_TYPE = 0, FORCE = 1, CNT_POINT = 2, COMPONENT = 4, NORMAL = 8, TORQUE = 16, CNT_WILDCARD = 32,
                          OWNER = 64, GEO_ID = 128, NICKNAME = 256 };
typedef bool objNormal_t;
const objNormal_t ENTITY_NORMAL_INWARD = 0;
const objNormal_t ENTITY_NORMAL_OUTWARD = 1;
constexpr unsigned int RESERVED_FAMILY_NUM = 255;

std::filesystem::path GET_DATA_PATH();
std::filesystem::path GetDEMEDataFile(const std::string& relative);
/// Tell the library where the reference-style data directory (clumps/, mesh/) lives
void SetDEMEDataPath(const std::string& path);

// ---------------------------------------------------------------------------------------------------------------
class DEMMaterial {
  public:
    std::unordered_map<std::string, float> mat_prop;
    unsigned int load_order = 0;
    DEMMaterial(const std::unordered_map<std::string, float>& prop) : mat_prop(prop) {}
};

class DEMInitializer {
  public:
    OWNER_TYPE obj_type = OWNER_TYPE::CLUMP;
    unsigned int load_order = 0;
    virtual ~DEMInitializer() {}
};

class DEMClumpTemplate {
  public:
    float mass = 0;
    float3 MOI = make_float3(0, 0, 0);
    std::vector<float> radii;
    std::vector<float3> relPos;
    std::vector<std::shared_ptr<DEMMaterial>> materials;
    unsigned int nComp = 0;
    unsigned int mark = 0;
    float volume = 0;
    std::string m_name = "NULL";

    int ReadComponentFromFile(const std::string filename, const std::string x_id = "x", const std::string y_id = "y",
                              const std::string z_id = "z", const std::string r_id = "r");
    void SetMass(float m) { mass = m; }
    void SetMOI(float3 moi) { MOI = moi; }
    void SetMaterial(const std::shared_ptr<DEMMaterial>& input) { materials.assign(nComp, input); }
    void SetVolume(float vol) { volume = vol; }
    void Scale(float s);
    /// The component positions were given in a frame that is not the centroid / principal frame: report where that
    /// frame sits (Structs.h:650-665 of the reference) and relPos is re-expressed in it ...
    void InformCentroidPrincipal(float3 center, float4 prin_Q) {
        for (auto& pos : relPos) applyFrameTransformGlobalToLocal(pos, center, prin_Q);
    }
    void InformCentroidPrincipal(const std::vector<float>& center, const std::vector<float>& prin_Q) {
        InformCentroidPrincipal(vec3_arg(center, "InformCentroidPrincipal"), vec4_arg(prin_Q, "InformCentroidPrincipal"));
    }
    /// ... or rotate, then move the components (:667-679)
    void Move(float3 vec, float4 rot_Q) {
        for (auto& pos : relPos) applyFrameTransformLocalToGlobal(pos, vec, rot_Q);
    }
    void Move(const std::vector<float>& vec, const std::vector<float>& rot_Q) { Move(vec3_arg(vec, "Move"), vec4_arg(rot_Q, "Move")); }
    void AssignName(const std::string& some_name) { m_name = some_name; }

  private:
    static float3 vec3_arg(const std::vector<float>& v, const char* who) {
        if (v.size() != 3) throw std::runtime_error(std::string(who) + ": a 3-element vector is expected");
        return make_float3(v[0], v[1], v[2]);
    }
    static float4 vec4_arg(const std::vector<float>& v, const char* who) {
        if (v.size() != 4) throw std::runtime_error(std::string(who) + ": a 4-element vector (x, y, z, w) is expected");
        return make_float4(v[0], v[1], v[2], v[3]);
    }
};

/// One facet, by value (src/DEM/Structs.h:550-558)
class DEMTriangle {
  public:
    DEMTriangle(float3 pnt1, float3 pnt2, float3 pnt3) : p1(pnt1), p2(pnt2), p3(pnt3) {}
    DEMTriangle() {}
    float3 p1, p2, p3;
};

class DEMClumpBatch : public DEMInitializer {
  public:
    size_t nClumps = 0;
    size_t nSpheres = 0;
    bool family_isSpecified = false;
    std::vector<std::shared_ptr<DEMClumpTemplate>> types;
    std::vector<unsigned int> families;
    std::vector<float3> vel, angVel, xyz;
    std::vector<float4> oriQ;  // x,y,z,w
    std::vector<std::pair<bodyID_t, bodyID_t>> contact_pairs;
    std::unordered_map<std::string, std::vector<float>> contact_wildcards;
    bodyID_t first_owner = 0;  // resolved at Initialize()

    explicit DEMClumpBatch(size_t num);
    size_t GetNumClumps() const { return nClumps; }
    size_t GetNumSpheres() const { return nSpheres; }
    void SetTypes(const std::vector<std::shared_ptr<DEMClumpTemplate>>& input);
    void SetTypes(const std::shared_ptr<DEMClumpTemplate>& input) { SetTypes(std::vector<std::shared_ptr<DEMClumpTemplate>>(nClumps, input)); }
    void SetType(const std::shared_ptr<DEMClumpTemplate>& input) { SetTypes(input); }
    void SetPos(const std::vector<float3>& input);
    void SetPos(float3 input) { SetPos(std::vector<float3>(nClumps, input)); }
    void SetVel(const std::vector<float3>& input);
    void SetVel(float3 input) { SetVel(std::vector<float3>(nClumps, input)); }
    void SetAngVel(const std::vector<float3>& input);
    void SetAngVel(float3 input) { SetAngVel(std::vector<float3>(nClumps, input)); }
    void SetOriQ(const std::vector<float4>& input);
    void SetOriQ(float4 input) { SetOriQ(std::vector<float4>(nClumps, input)); }
    // {x, y, z} (one value for all clumps) and {{x, y, z}, ...} forms (Structs.h:763-830 of the reference)
    void SetPos(const std::vector<float>& input) { SetPos(vec3_of(input, "SetPos")); }
    void SetPos(const std::vector<std::vector<float>>& input) { SetPos(vec3s_of(input, "SetPos")); }
    void SetVel(const std::vector<float>& input) { SetVel(vec3_of(input, "SetVel")); }
    void SetVel(const std::vector<std::vector<float>>& input) { SetVel(vec3s_of(input, "SetVel")); }
    void SetAngVel(const std::vector<float>& input) { SetAngVel(vec3_of(input, "SetAngVel")); }
    void SetAngVel(const std::vector<std::vector<float>>& input) { SetAngVel(vec3s_of(input, "SetAngVel")); }
    void SetOriQ(const std::vector<float>& input) { SetOriQ(vec4_of(input, "SetOriQ")); }
    void SetOriQ(const std::vector<std::vector<float>>& input) {
        std::vector<float4> q;
        for (const auto& v : input) q.push_back(vec4_of(v, "SetOriQ"));
        SetOriQ(q);
    }
    void SetFamilies(const std::vector<unsigned int>& input);
    void SetFamilies(unsigned int input) { SetFamilies(std::vector<unsigned int>(nClumps, input)); }
    void SetFamily(unsigned int input) { SetFamilies(input); }
    /// Restart support: contacts (pairs of sphere numbers relative to this batch) and their history
    void SetExistingContacts(const std::vector<std::pair<bodyID_t, bodyID_t>>& pairs) { contact_pairs = pairs; }
    void SetExistingContactWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards) { contact_wildcards = wildcards; }
    void AddExistingContactWildcard(const std::string& name, const std::vector<float>& vals);
    /// Owner / geometry wildcards exist only for custom force models (which need run-time compilation and are not part
    /// of this core): the values are kept, and Initialize() refuses a batch that carries any, as the reference refuses a
    /// wildcard its force model does not declare (Structs.h:868-917, APIPrivate.cpp:1158-1190).
    void SetOwnerWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards);
    void AddOwnerWildcard(const std::string& name, const std::vector<float>& vals);
    void AddOwnerWildcard(const std::string& name, float val) { AddOwnerWildcard(name, std::vector<float>(nClumps, val)); }
    void SetGeometryWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards);
    void AddGeometryWildcard(const std::string& name, const std::vector<float>& vals);
    void AddGeometryWildcard(const std::string& name, float val) { AddGeometryWildcard(name, std::vector<float>(nSpheres, val)); }
    std::unordered_map<std::string, std::vector<float>> owner_wildcards, geo_wildcards;
    size_t GetNumContacts() const { return contact_pairs.size(); }

  private:
    void assertLength(size_t len, const std::string& name) const;
    static float3 vec3_of(const std::vector<float>& v, const char* who) {
        if (v.size() != 3) throw std::runtime_error(std::string(who) + ": a 3-element vector is expected");
        return make_float3(v[0], v[1], v[2]);
    }
    static float4 vec4_of(const std::vector<float>& v, const char* who) {
        if (v.size() != 4) throw std::runtime_error(std::string(who) + ": a 4-element vector (x, y, z, w) is expected");
        return make_float4(v[0], v[1], v[2], v[3]);
    }
    static std::vector<float3> vec3s_of(const std::vector<std::vector<float>>& in, const char* who) {
        std::vector<float3> out;
        for (const auto& v : in) out.push_back(vec3_of(v, who));
        return out;
    }
};

class DEMExternObj : public DEMInitializer {
  public:
    struct Component {
        int type;  // 0 plane, 2 infinite cylinder
        float3 pos, dir;
        float size1;
        float normal;
        std::shared_ptr<DEMMaterial> material;
    };
    std::vector<Component> comps;
    unsigned int family_code = RESERVED_FAMILY_NUM;
    float3 init_pos = make_float3(0, 0, 0);
    float4 init_oriQ = make_float4(0, 0, 0, 1);
    float mass = 1e6;
    float3 MOI = make_float3(1e6, 1e6, 1e6);
    bodyID_t owner = 0;  // resolved at Initialize()

    DEMExternObj() { obj_type = OWNER_TYPE::ANALYTICAL; }
    void SetFamily(const unsigned int code);
    void SetMass(float m) { mass = m; }
    void SetMOI(float3 moi) { MOI = moi; }
    void SetInitQuat(const float4 rotQ) { init_oriQ = rotQ; }
    void SetInitPos(const float3 displ) { init_pos = displ; }
    void AddPlane(const float3 pos, const float3 normal, const std::shared_ptr<DEMMaterial>& material);
    void AddZCylinder(const float3 pos, const float rad, const std::shared_ptr<DEMMaterial>& material,
                      const objNormal_t normal = ENTITY_NORMAL_INWARD);
    void AddCylinder(const float3 pos, const float3 axis, const float rad, const std::shared_ptr<DEMMaterial>& material,
                     const objNormal_t normal = ENTITY_NORMAL_INWARD);
    // std::vector<float> forms of the above (BdrsAndObjs.h:118-228 of the reference).  The reference's finite plate
    // (AddPlate, :162-177) is commented out there and its device case never reports a contact
    // (DEMHelperKernels.cuh:491-493), so there is no plate to mirror.
    void SetMOI(const std::vector<float>& moi) { SetMOI(three(moi, "SetMOI")); }
    void SetInitPos(const std::vector<float>& displ) { SetInitPos(three(displ, "SetInitPos")); }
    void SetInitQuat(const std::vector<float>& rotQ) {
        if (rotQ.size() != 4) throw std::runtime_error("SetInitQuat: a 4-element vector (x, y, z, w) is expected");
        SetInitQuat(make_float4(rotQ[0], rotQ[1], rotQ[2], rotQ[3]));
    }
    void AddPlane(const std::vector<float>& pos, const std::vector<float>& normal, const std::shared_ptr<DEMMaterial>& material) {
        AddPlane(three(pos, "AddPlane"), three(normal, "AddPlane"), material);
    }
    void AddZCylinder(const std::vector<float>& pos, const float rad, const std::shared_ptr<DEMMaterial>& material,
                      const objNormal_t normal = ENTITY_NORMAL_INWARD) {
        AddZCylinder(three(pos, "AddZCylinder"), rad, material, normal);
    }
    void AddCylinder(const std::vector<float>& pos, const std::vector<float>& axis, const float rad,
                     const std::shared_ptr<DEMMaterial>& material, const objNormal_t normal = ENTITY_NORMAL_INWARD) {
        AddCylinder(three(pos, "AddCylinder"), three(axis, "AddCylinder"), rad, material, normal);
    }

  private:
    static float3 three(const std::vector<float>& v, const char* who) {
        if (v.size() != 3) throw std::runtime_error(std::string(who) + ": a 3-element vector is expected");
        return make_float3(v[0], v[1], v[2]);
    }
};

/// Triangle mesh owner (src/DEM/BdrsAndObjs.h:222-520). Vertices live in the mesh frame; facets are counter-clockwise
/// (the right-hand-rule normal is the side that pushes spheres away).
class DEMMeshConnected : public DEMInitializer {
  public:
    size_t nTri = 0;
    size_t tri_first = 0;  // id of this mesh's first facet in the flattened facet arrays (set by Initialize)
    std::vector<float3> m_vertices;
    std::vector<float3> m_normals;
    std::vector<float3> m_UV;
    std::vector<int3> m_face_v_indices;
    std::vector<int3> m_face_n_indices;
    std::vector<int3> m_face_uv_indices;
    std::vector<float3> m_colors;
    std::vector<int3> m_face_col_indices;
    bool use_mesh_normals = false;
    std::vector<std::shared_ptr<DEMMaterial>> materials;
    bool isMaterialSet = false;
    unsigned int family_code = RESERVED_FAMILY_NUM;
    float3 init_pos = make_float3(0, 0, 0);
    float4 init_oriQ = make_float4(0, 0, 0, 1);
    float mass = 1.f;
    float3 MOI = make_float3(1.f, 1.f, 1.f);
    std::string filename;
    bodyID_t owner = 0;  // resolved at Initialize()

    DEMMeshConnected() { obj_type = OWNER_TYPE::MESH; }
    explicit DEMMeshConnected(const std::string& input_file) {
        obj_type = OWNER_TYPE::MESH;
        LoadWavefrontMesh(input_file);
    }
    DEMMeshConnected(const std::string& input_file, const std::shared_ptr<DEMMaterial>& mat) {
        obj_type = OWNER_TYPE::MESH;
        LoadWavefrontMesh(input_file);
        SetMaterial(mat);
    }
    bool LoadWavefrontMesh(const std::string& input_file, bool load_normals = true, bool load_uv = false);
    /// Build from raw arrays (vertices + vertex-index triples)
    void SetGeometry(const std::vector<float3>& vertices, const std::vector<int3>& faces);
    size_t GetNumTriangles() const { return nTri; }
    size_t GetNumNodes() const { return m_vertices.size(); }
    std::vector<float3>& GetCoordsVertices() { return m_vertices; }
    std::vector<std::vector<float>> GetCoordsVerticesAsVectorOfVectors();
    std::vector<float3>& GetCoordsNormals() { return m_normals; }
    std::vector<float3>& GetCoordsUV() { return m_UV; }
    std::vector<float3>& GetCoordsColors() { return m_colors; }
    std::vector<int3>& GetIndicesVertexes() { return m_face_v_indices; }
    std::vector<std::vector<int>> GetIndicesVertexesAsVectorOfVectors();
    std::vector<int3>& GetIndicesNormals() { return m_face_n_indices; }
    std::vector<int3>& GetIndicesUV() { return m_face_uv_indices; }
    std::vector<int3>& GetIndicesColors() { return m_face_col_indices; }
    /// the facet normals always come from the winding of the nodes here (as in the reference's contact kernels); the flag
    /// is kept for scripts that set it
    void UseNormals(bool use = true) { use_mesh_normals = use; }
    DEMTriangle GetTriangle(size_t index) const {
        const int3 f = m_face_v_indices.at(index);
        return DEMTriangle(m_vertices[f.x], m_vertices[f.y], m_vertices[f.z]);
    }
    /// All meshes as ONE Wavefront object: vertices, normals where a mesh has them, faces (src/DEM/MeshUtils.cpp:137-184)
    static void WriteWavefront(const std::string& filename, std::vector<DEMMeshConnected>& meshes);
    /// One mesh holding the nodes, normals, facets and materials of all (the reference declares this without defining it)
    static DEMMeshConnected Merge(std::vector<DEMMeshConnected>& meshes);
    // per-facet wildcards exist only for custom force models: kept, and refused at Initialize() like a batch's
    std::unordered_map<std::string, std::vector<float>> geo_wildcards;
    void ClearWildcards() { geo_wildcards.clear(); }
    void SetGeometryWildcards(const std::unordered_map<std::string, std::vector<float>>& wildcards);
    void AddGeometryWildcard(const std::string& name, const std::vector<float>& vals);
    void AddGeometryWildcard(const std::string& name, float val) { AddGeometryWildcard(name, std::vector<float>(nTri, val)); }
    void Clear();
    void SetMass(float m) { mass = m; }
    void SetMOI(float3 moi) { MOI = moi; }
    void SetFamily(unsigned int num) { family_code = num; }
    void SetMaterial(const std::vector<std::shared_ptr<DEMMaterial>>& input);
    void SetMaterial(const std::shared_ptr<DEMMaterial>& input) { SetMaterial(std::vector<std::shared_ptr<DEMMaterial>>(nTri, input)); }
    void SetInitQuat(const float4 rotQ) { init_oriQ = rotQ; }
    void SetInitPos(const float3 displ) { init_pos = displ; }
    void InformCentroidPrincipal(float3 center, float4 prin_Q);
    void Move(float3 vec, float4 rot_Q);
    void Mirror(float3 plane_point, float3 plane_normal);
    void Scale(float s);
    void Scale(float3 s);
};

class DEMSolver;

/// Per-contact read-out of GetContactDetailedInfo (src/DEM/Structs.h:1049-1107): one entry per reported contact in
/// every field that SetContactOutputContent switched on; asking for a field that is off throws.
class ContactInfoContainer {
  public:
    ContactInfoContainer(unsigned int cnt_out_content, const std::vector<std::string>& wildcard_names);
    size_t Size() const { return m_type.size(); }
    std::vector<std::string>& GetContactType() { return m_type; }
    std::vector<float3>& GetPoint() { return need(m_point, CNT_POINT, "Point"); }
    std::vector<bodyID_t>& GetAOwner() { return need(m_owner[0], OWNER, "AOwner"); }
    std::vector<bodyID_t>& GetBOwner() { return need(m_owner[1], OWNER, "BOwner"); }
    std::vector<bodyID_t>& GetAGeo() { return need(m_geo[0], GEO_ID, "AGeo"); }
    std::vector<bodyID_t>& GetBGeo() { return need(m_geo[1], GEO_ID, "BGeo"); }
    std::vector<family_t>& GetAOwnerFamily() { return m_family[0]; }
    std::vector<family_t>& GetBOwnerFamily() { return m_family[1]; }
    std::vector<float3>& GetForce() { return need(m_force, FORCE, "Force"); }
    std::vector<float3>& GetTorque() { return need(m_torque, TORQUE, "Torque"); }
    std::vector<float3>& GetNormal() { return need(m_normal, NORMAL, "Normal"); }
    /// a contact wildcard by name (delta_tan_x, delta_tan_y, delta_tan_z, delta_time for the frictional model)
    std::vector<float>& GetWildcard(const std::string& name);
    const std::vector<std::string>& GetWildcardNames() const { return m_wc_names; }
    bool Contains(unsigned int content_bit) const { return (m_content & content_bit) != 0; }
    void ResizeAll(size_t n);

  private:
    template <typename V>
    V& need(V& field, unsigned int bit, const char* key) {
        if (!(m_content & bit))
            throw std::runtime_error(std::string("ContactInfoContainer does not have field: '") + key +
                                     "', you may need to turn on the output of this field by correctly calling "
                                     "SetContactOutputContent before Initialize().");
        return field;
    }
    unsigned int m_content;
    std::vector<std::string> m_type, m_wc_names;
    std::vector<float3> m_point, m_force, m_torque, m_normal;
    std::vector<bodyID_t> m_owner[2], m_geo[2];
    std::vector<family_t> m_family[2];
    std::vector<std::vector<float>> m_wc;
};

/// Tracker of one loaded object (a clump batch, an external object or a mesh): src/DEM/AuxClasses.h:93-420.  Offsets
/// count owners from the first one of the tracked object; the plural getters read the whole object in one transfer.
class DEMTracker {
  public:
    DEMTracker(DEMSolver* sim, std::shared_ptr<DEMInitializer> obj) : sys(sim), obj(std::move(obj)) {}
    bodyID_t GetOwnerID(size_t offset = 0);
    std::vector<bodyID_t> GetOwnerIDs();
    float3 Pos(size_t offset = 0);
    std::vector<float> GetPos(size_t offset = 0) { return Real3ToVec(Pos(offset)); }
    std::vector<float3> Positions();
    std::vector<std::vector<float>> GetPositions() { return Real3VectorToVecOfVec(Positions()); }
    float3 Vel(size_t offset = 0);
    std::vector<float> GetVel(size_t offset = 0) { return Real3ToVec(Vel(offset)); }
    std::vector<float3> Velocities();
    std::vector<std::vector<float>> GetVelocities() { return Real3VectorToVecOfVec(Velocities()); }
    float3 AngVelLocal(size_t offset = 0);
    std::vector<float> GetAngVelLocal(size_t offset = 0) { return Real3ToVec(AngVelLocal(offset)); }
    std::vector<float3> AngularVelocitiesLocal();
    std::vector<std::vector<float>> GetAngularVelocitiesLocal() { return Real3VectorToVecOfVec(AngularVelocitiesLocal()); }
    float3 AngVelGlobal(size_t offset = 0);
    std::vector<float> GetAngVelGlobal(size_t offset = 0) { return Real3ToVec(AngVelGlobal(offset)); }
    std::vector<float3> AngularVelocitiesGlobal();
    std::vector<std::vector<float>> GetAngularVelocitiesGlobal() { return Real3VectorToVecOfVec(AngularVelocitiesGlobal()); }
    float4 OriQ(size_t offset = 0);
    std::vector<float> GetOriQ(size_t offset = 0) { return Real4ToVec(OriQ(offset)); }
    std::vector<float4> OrientationQuaternions();
    std::vector<std::vector<float>> GetOrientationQuaternions() { return Real4VectorToVecOfVec(OrientationQuaternions()); }
    unsigned int GetFamily(size_t offset = 0);
    std::vector<unsigned int> GetFamilies();
    /// clumps in (potential) contact with the tracked owner at `offset`
    std::vector<bodyID_t> GetContactClumps(size_t offset = 0);
    /// acceleration / angular acceleration that the contacts of the last step gave the owner
    float3 ContactAcc(size_t offset = 0);
    std::vector<float> GetContactAcc(size_t offset = 0) { return Real3ToVec(ContactAcc(offset)); }
    std::vector<float3> ContactAccelerations();
    std::vector<std::vector<float>> GetContactAccelerations() { return Real3VectorToVecOfVec(ContactAccelerations()); }
    float3 ContactAngAccLocal(size_t offset = 0);
    std::vector<float> GetContactAngAccLocal(size_t offset = 0) { return Real3ToVec(ContactAngAccLocal(offset)); }
    std::vector<float3> ContactAngularAccelerationsLocal();
    std::vector<std::vector<float>> GetContactAngularAccelerationsLocal() {
        return Real3VectorToVecOfVec(ContactAngularAccelerationsLocal());
    }
    float3 ContactAngAccGlobal(size_t offset = 0);
    std::vector<float> GetContactAngAccGlobal(size_t offset = 0) { return Real3ToVec(ContactAngAccGlobal(offset)); }
    std::vector<float3> ContactAngularAccelerationsGlobal();
    std::vector<std::vector<float>> GetContactAngularAccelerationsGlobal() {
        return Real3VectorToVecOfVec(ContactAngularAccelerationsGlobal());
    }
    float Mass(size_t offset = 0);
    std::vector<float> Masses();
    float3 MOI(size_t offset = 0);
    std::vector<float> GetMOI(size_t offset = 0) { return Real3ToVec(MOI(offset)); }
    std::vector<float3> MOIs();
    std::vector<std::vector<float>> GetMOIs() { return Real3VectorToVecOfVec(MOIs()); }
    /// Owner / geometry wildcards belong to custom force models; the built-in models declare none, so -- like the
    /// reference for a name its force model does not know -- these throw.
    float GetOwnerWildcardValue(const std::string& name, size_t offset = 0);
    std::vector<float> GetOwnerWildcardValues(const std::string& name);
    float GetGeometryWildcardValue(const std::string& name, size_t offset);
    std::vector<float> GetGeometryWildcardValues(const std::string& name);
    void SetOwnerWildcardValue(const std::string& name, float wc, size_t offset = 0);
    void SetOwnerWildcardValues(const std::string& name, const std::vector<float>& wc);
    void SetGeometryWildcardValue(const std::string& name, float wc, size_t offset = 0);
    void SetGeometryWildcardValues(const std::string& name, const std::vector<float>& wc);
    /// Deforming mesh (AuxClasses.h:288-313 of the reference): replace / displace the tracked mesh's nodes (mesh frame;
    /// one entry per node), read the nodes back in the global frame, get the mesh handle
    void UpdateMesh(const std::vector<float3>& new_nodes);
    void UpdateMeshByIncrement(const std::vector<float3>& deformation);
    std::vector<float3> GetMeshNodesGlobal();
    std::shared_ptr<DEMMeshConnected> GetMesh();
    /// Contact points and forces (global frame) acting on the tracked owner at `offset` / on all tracked owners
    size_t GetContactForces(std::vector<float3>& points, std::vector<float3>& forces, size_t offset = 0);
    size_t GetContactForcesForAll(std::vector<float3>& points, std::vector<float3>& forces);
    /// ... plus the torque of each contact that is not already the moment of its force: the rolling-resistance couple,
    /// about the owner's centre, in the global or in the owner's frame (AuxClasses.h:372-418)
    size_t GetContactForcesAndGlobalTorque(std::vector<float3>& points, std::vector<float3>& forces,
                                           std::vector<float3>& torques, size_t offset = 0);
    size_t GetContactForcesAndGlobalTorqueForAll(std::vector<float3>& points, std::vector<float3>& forces,
                                                 std::vector<float3>& torques);
    size_t GetContactForcesAndLocalTorque(std::vector<float3>& points, std::vector<float3>& forces,
                                          std::vector<float3>& torques, size_t offset = 0);
    size_t GetContactForcesAndLocalTorqueForAll(std::vector<float3>& points, std::vector<float3>& forces,
                                                std::vector<float3>& torques);
    void SetPos(float3 pos, size_t offset = 0);
    void SetPos(const std::vector<float3>& pos);
    void SetVel(float3 vel, size_t offset = 0);
    void SetVel(const std::vector<float3>& vel);
    void SetAngVel(float3 angVel, size_t offset = 0);
    void SetAngVel(const std::vector<float3>& angVel);
    void SetOriQ(float4 oriQ, size_t offset = 0);
    void SetOriQ(const std::vector<float4>& oriQ);
    /// Extra (angular) acceleration for the NEXT time step only, added to what the contacts give (AuxClasses.h:262-274)
    void AddAcc(float3 acc, size_t offset = 0);
    void AddAcc(const std::vector<float3>& acc);
    void AddAngAcc(float3 angAcc, size_t offset = 0);
    void AddAngAcc(const std::vector<float3>& angAcc);
    void SetFamily(unsigned int fam_num);
    void SetFamily(unsigned int fam_num, size_t offset);
    void ChangeClumpSizes(const std::vector<bodyID_t>& IDs, const std::vector<float>& factors);

  private:
    DEMSolver* sys;
    std::shared_ptr<DEMInitializer> obj;
    bodyID_t first();
    size_t count();
    void assertOwnerSize(size_t input_length, const std::string& name);
};

/// Built-in inspectors (src/DEM/AuxClasses.cpp:88-164): clump_max_z, clump_min_z, clump_max_absv, max_absv,
/// clump_kinetic_energy, clump_mass, clump_volume.  Over the whole domain they are device reductions; with a region -- a
/// condition on the position X, Y, Z of the inspected thing (the sphere for the *_z and clump_max_absv quantities, the
/// owner otherwise), e.g. "return (abs(X) <= 0.48) && (Z <= -0.44);" -- the owners are read back and the condition is
/// evaluated on the host (DEM/utils/Expression.hpp; the reference compiles it into its inspection kernel).
class DEMInspector {
  public:
    DEMInspector(DEMSolver* sim, const std::string& quantity);
    DEMInspector(DEMSolver* sim, const std::string& quantity, const std::string& region);
    float GetValue();
    /// The un-reduced form: one value per owner for the quantity "absv" (every owner's speed), in owner order.  The
    /// pointer stays valid until the next call.
    float* GetValues();
    /// Inspection code is C++ text the reference compiles into its query kernel: not available here
    void SetInspectionCode(const std::string& code);

  private:
    DEMSolver* sys;
    int kind;
    std::shared_ptr<ScalarExpression> region;
    std::vector<float> m_values;
};

/// The force model handle returned by Use*Model / DefineContactForceModel (src/DEM/AuxClasses.h:424-520).  The two
/// built-in models are compiled into the core; their material requirements and history words are fixed.
class DEMForceModel {
  public:
    explicit DEMForceModel(FORCE_MODEL t) : type(t) {}
    FORCE_MODEL type;
    /// material properties every material must define / that may be set pairwise: checked when materials are flattened
    void SetMustHaveMatProp(const std::set<std::string>& props) { m_must_have_mat_props = props; }
    void SetMustPairwiseMatProp(const std::set<std::string>& props) { m_pairwise_mat_props = props; }
    /// History words and owner / geometry wildcards are what a custom model declares for its own code; the built-in
    /// models keep theirs (delta_tan_x/y/z, delta_time for the frictional model, none for the frictionless one).
    void SetPerContactWildcards(const std::set<std::string>& wildcards);
    void SetPerOwnerWildcards(const std::set<std::string>& wildcards);
    void SetPerGeometryWildcards(const std::set<std::string>& wildcards);
    void SetForceModelType(FORCE_MODEL model_type);
    void DefineCustomModel(const std::string& model);
    int ReadCustomModelFile(const std::filesystem::path& sourcefile);
    void DefineCustomModelPrerequisites(const std::string& util);
    int ReadCustomModelPrerequisitesFile(const std::filesystem::path& sourcefile);
    std::set<std::string> m_must_have_mat_props, m_pairwise_mat_props;
};

// ---------------------------------------------------------------------------------------------------------------
class DEMSolver {
  public:
    /// nGPUs devices (those the box has, at most 8) form one group; large scenes are sharded over them (x-slabs)
    explicit DEMSolver(unsigned int nGPUs = 2);
    explicit DEMSolver(const std::vector<int>& gpu_ids);
    ~DEMSolver();
    DEMSolver(const DEMSolver&) = delete;
    DEMSolver& operator=(const DEMSolver&) = delete;

    void SetVerbosity(VERBOSITY verbose) { verbosity = verbose; }
    void SetVerbosity(const std::string& verbose);
    /// Files are written as CSV.  As in a reference build without ChPF (APIPublic.cpp:171-210): "CHPF" is refused, "BINARY"
    /// is accepted and answered with CSV plus a warning at write time, anything else is an error.
    void SetOutputFormat(OUTPUT_FORMAT format);
    void SetOutputFormat(const std::string& format);
    void SetOutputContent(unsigned int content) { m_out_content = content; }
    void SetOutputContent(const std::vector<std::string>& content);
    void SetContactOutputContent(unsigned int content) { m_cnt_out_content = content; }
    void SetContactOutputContent(const std::vector<std::string>& content);
    /// Meshes are written as VTK; OBJ is accepted here and refused by WriteMeshFile, as in the reference (:2082-2096)
    void SetMeshOutputFormat(MESH_FORMAT format) { m_mesh_out_format = format; }
    void SetMeshOutputFormat(const std::string& format);

    void InstructBoxDomainDimension(float x, float y, float z, const std::string& dir_exact = "none");
    void InstructBoxDomainDimension(const std::pair<float, float>& x, const std::pair<float, float>& y,
                                    const std::pair<float, float>& z, const std::string& dir_exact = "none");
    void InstructBoxDomainBoundingBC(const std::string& inst, const std::shared_ptr<DEMMaterial>& mat);
    void SetGravitationalAcceleration(float3 g) { G = g; }
    void SetInitTimeStep(double ts_size) { m_ts_size = ts_size; }
    void UpdateStepSize(double ts = -1.0);
    double GetTimeStepSize() const { return m_ts_size; }
    size_t GetNumClumps() const { return nOwnerClumps; }
    size_t GetNumOwners() const { return nOwnerBodies; }
    size_t GetNumContacts() const;
    float GetAvgSphContacts() const;
    double GetSimTime() const;
    void SetSimTime(double time);
    /// Broad-phase cell size and cell count of the last contact-list rebuild (the reference's bin size / number of bins)
    double GetBinSize() const;
    size_t GetBinNum() const;
    /// The contact margin currently added to every sphere radius (largest over the owners)
    float GetExpandFactor() const;
    size_t GetDeviceMemUsageDynamic() const;
    size_t GetDeviceMemUsageKinematic() const { return 0; }  // one context does both jobs here
    /// Owner ids of the clumps in (potential) contact with the given owner
    std::vector<bodyID_t> GetOwnerContactClumps(bodyID_t ownerID) const;
    float GetUpdateFreq() const;
    bool GetInitStatus() const { return sys_initialized; }

    void SetCDUpdateFreq(int freq);
    void SetIntegrator(const std::string& intg);
    void SetIntegrator(TIME_INTEGRATOR intg) { m_integrator = intg; }
    void SetExpandFactor(float beta, bool fix = true);
    void SetMaxVelocity(float max_vel) { m_approx_max_vel = max_vel; }
    /// only "auto" exists (margin from the measured top speed, the default): anything else is an error, as in the reference
    void SetExpandSafetyType(const std::string& insp_type);
    void SetExpandSafetyMultiplier(float param) { m_expand_safety_multi = param; }
    void SetExpandSafetyAdder(float vel) { m_expand_base_vel = vel; }
    void SetErrorOutVelocity(float vel) { threshold_error_out_vel = vel; }
    /// stop with an error when a sphere has more listed contacts than this on average (default 100, Structs.h:527)
    void SetErrorOutAvgContacts(float num_cnts) { threshold_error_out_num_cnts = num_cnts; }
    void SetNoForceRecord(bool flag = true) { no_recording_contact_forces = flag; }
    // Tuning knobs of the reference's two-thread / jitify machinery: accepted and ignored
    void SetInitBinSize(double) {}
    void SetInitBinSizeAsMultipleOfSmallestSphere(float) {}
    void SetInitBinNumTarget(size_t) {}
    void InstructNumOwners(size_t) {}
    void SetSortContactPairs(bool) {}
    void SetJitifyClumpTemplates(bool = true) {}
    void DisableJitifyClumpTemplates() {}
    void SetJitifyMassProperties(bool = true) {}
    void DisableJitifyMassProperties() {}
    void UseCompactForceKernel(bool) {}
    void UseCubForceCollection(bool = true) {}
    void SetCollectAccRightAfterForceCalc(bool = true) {}
    void UseAdaptiveBinSize(bool = true) {}
    void DisableAdaptiveBinSize() {}
    /// let the solver pick the steps per contact-list cycle (on by default, as in the reference; SetCDUpdateFreq gives the
    /// starting point, SetCDMaxUpdateFreq the upper bound)
    void UseAdaptiveUpdateFreq(bool flag = true);
    void DisableAdaptiveUpdateFreq() { UseAdaptiveUpdateFreq(false); }
    void SetAdaptiveBinSizeDelaySteps(unsigned int) {}
    void SetAdaptiveBinSizeMaxRate(float) {}
    void SetAdaptiveBinSizeAcc(float) {}
    void SetAdaptiveBinSizeUpperProactivity(float) {}
    void SetAdaptiveBinSizeLowerProactivity(float) {}
    void SetCDMaxUpdateFreq(unsigned int max_freq);
    void SetCDNumStepsMaxDriftAheadOfAvg(float) {}
    void SetCDNumStepsMaxDriftMultipleOfAvg(float) {}
    void SetCDNumStepsMaxDriftHistorySize(unsigned int) {}
    void SetMaxSphereInBin(unsigned int) {}
    void SetMaxTriangleInBin(unsigned int) {}
    void SetForceCalcThreadsPerBlock(unsigned int) {}

    std::shared_ptr<DEMForceModel> UseFrictionalHertzianModel();
    std::shared_ptr<DEMForceModel> UseFrictionlessHertzianModel();
    std::shared_ptr<DEMForceModel> DefineContactForceModel(const std::string&);
    std::shared_ptr<DEMForceModel> ReadContactForceModel(const std::string&);
    std::shared_ptr<DEMForceModel> GetContactForceModel() { return m_force_model_obj; }

    std::shared_ptr<DEMMaterial> LoadMaterial(const std::unordered_map<std::string, float>& mat_prop);
    std::shared_ptr<DEMMaterial> LoadMaterial(DEMMaterial& a_material) { return LoadMaterial(a_material.mat_prop); }
    // deep copies of an already loaded material / clump template / batch, loaded as new objects
    // (reference src/DEM/APIPublic.cpp:397-411)
    std::shared_ptr<DEMMaterial> Duplicate(const std::shared_ptr<DEMMaterial>& ptr) {
        DEMMaterial obj = *ptr;
        return LoadMaterial(obj);
    }
    std::shared_ptr<DEMClumpTemplate> Duplicate(const std::shared_ptr<DEMClumpTemplate>& ptr) {
        DEMClumpTemplate obj = *ptr;
        return LoadClumpType(obj);
    }
    std::shared_ptr<DEMClumpBatch> Duplicate(const std::shared_ptr<DEMClumpBatch>& ptr) {
        DEMClumpBatch obj = *ptr;
        return AddClumps(obj);
    }
    void SetMaterialPropertyPair(const std::string& name, const std::shared_ptr<DEMMaterial>& mat1,
                                 const std::shared_ptr<DEMMaterial>& mat2, float val);
    std::shared_ptr<DEMClumpTemplate> LoadClumpType(float mass, float3 moi, const std::vector<float>& sp_radii,
                                                    const std::vector<float3>& sp_locations_xyz,
                                                    const std::vector<std::shared_ptr<DEMMaterial>>& sp_materials);
    std::shared_ptr<DEMClumpTemplate> LoadClumpType(float mass, float3 moi, const std::vector<float>& sp_radii,
                                                    const std::vector<float3>& sp_locations_xyz,
                                                    const std::shared_ptr<DEMMaterial>& sp_material);
    std::shared_ptr<DEMClumpTemplate> LoadClumpType(float mass, float3 moi, const std::string filename,
                                                    const std::vector<std::shared_ptr<DEMMaterial>>& sp_materials);
    std::shared_ptr<DEMClumpTemplate> LoadClumpType(float mass, float3 moi, const std::string filename,
                                                    const std::shared_ptr<DEMMaterial>& sp_material);
    std::shared_ptr<DEMClumpTemplate> LoadClumpType(DEMClumpTemplate& clump);
    std::shared_ptr<DEMClumpTemplate> LoadSphereType(float mass, float radius, const std::shared_ptr<DEMMaterial>& material);

    std::shared_ptr<DEMClumpBatch> AddClumps(DEMClumpBatch& input_batch);
    std::shared_ptr<DEMClumpBatch> AddClumps(const std::vector<std::shared_ptr<DEMClumpTemplate>>& input_types,
                                             const std::vector<float3>& input_xyz);
    std::shared_ptr<DEMClumpBatch> AddClumps(std::shared_ptr<DEMClumpTemplate>& input_type, float3 input_xyz) {
        return AddClumps(std::vector<std::shared_ptr<DEMClumpTemplate>>(1, input_type), std::vector<float3>(1, input_xyz));
    }
    std::shared_ptr<DEMClumpBatch> AddClumps(std::shared_ptr<DEMClumpTemplate>& input_type, const std::vector<float3>& input_xyz) {
        return AddClumps(std::vector<std::shared_ptr<DEMClumpTemplate>>(input_xyz.size(), input_type), input_xyz);
    }
    std::shared_ptr<DEMExternObj> AddExternalObject();
    std::shared_ptr<DEMExternObj> AddBCPlane(const float3 pos, const float3 normal, const std::shared_ptr<DEMMaterial>& material);
    std::shared_ptr<DEMMeshConnected> AddWavefrontMeshObject(const std::string& filename, const std::shared_ptr<DEMMaterial>& mat,
                                                             bool load_normals = true, bool load_uv = false);
    std::shared_ptr<DEMMeshConnected> AddWavefrontMeshObject(const std::string& filename, bool load_normals = true,
                                                             bool load_uv = false);
    std::shared_ptr<DEMMeshConnected> AddWavefrontMeshObject(DEMMeshConnected& mesh);
    std::shared_ptr<DEMMeshConnected> AddMesh(DEMMeshConnected& mesh) { return AddWavefrontMeshObject(mesh); }
    size_t GetNumMeshes() const { return m_cached_meshes.size(); }

    template <typename T>
    std::shared_ptr<DEMTracker> Track(const std::shared_ptr<T>& obj) {
        auto tr = std::make_shared<DEMTracker>(this, std::static_pointer_cast<DEMInitializer>(obj));
        m_trackers.push_back(tr);
        return tr;
    }
    std::shared_ptr<DEMInspector> CreateInspector(const std::string& quantity = "clump_max_z");
    std::shared_ptr<DEMInspector> CreateInspector(const std::string& quantity, const std::string& region);

    void DisableContactBetweenFamilies(unsigned int ID1, unsigned int ID2);
    void EnableContactBetweenFamilies(unsigned int ID1, unsigned int ID2);
    void SetFamilyFixed(unsigned int ID);
    /// all clumps (meshes) of family N take the material (after Initialize; src/DEM/API.h, APIPublic.cpp:1597-1604)
    void SetFamilyClumpMaterial(unsigned int N, const std::shared_ptr<DEMMaterial>& mat);
    void SetFamilyMeshMaterial(unsigned int N, const std::shared_ptr<DEMMaterial>& mat);
    /// (the reference re-compiles its kernels with line numbers in error messages; nothing is compiled at run time here)
    void EnsureKernelErrMsgLineNum(bool flag = true) { (void)flag; }
    void SetFamilyPrescribedLinVel(unsigned int ID, const std::string& velX, const std::string& velY,
                                   const std::string& velZ, bool dictate = true);
    void SetFamilyPrescribedAngVel(unsigned int ID, const std::string& velX, const std::string& velY,
                                   const std::string& velZ, bool dictate = true);
    void SetFamilyPrescribedPosition(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z,
                                     bool dictate = true);
    // "dictated, value left to the user" forms (API.h:712-786 of the reference): the listed components are no longer
    // influenced by contact forces; the user sets them through trackers
    void SetFamilyPrescribedLinVel(unsigned int ID) { markPrescribed(ID, 0, 7); }
    void SetFamilyPrescribedLinVelX(unsigned int ID) { markPrescribed(ID, 0, 1); }
    void SetFamilyPrescribedLinVelY(unsigned int ID) { markPrescribed(ID, 0, 2); }
    void SetFamilyPrescribedLinVelZ(unsigned int ID) { markPrescribed(ID, 0, 4); }
    void SetFamilyPrescribedAngVel(unsigned int ID) { markPrescribed(ID, 1, 7); }
    void SetFamilyPrescribedAngVelX(unsigned int ID) { markPrescribed(ID, 1, 1); }
    void SetFamilyPrescribedAngVelY(unsigned int ID) { markPrescribed(ID, 1, 2); }
    void SetFamilyPrescribedAngVelZ(unsigned int ID) { markPrescribed(ID, 1, 4); }
    void SetFamilyPrescribedPosition(unsigned int ID) { markPrescribed(ID, 2, 7); }
    void SetFamilyPrescribedPositionX(unsigned int ID) { markPrescribed(ID, 2, 1); }
    void SetFamilyPrescribedPositionY(unsigned int ID) { markPrescribed(ID, 2, 2); }
    void SetFamilyPrescribedPositionZ(unsigned int ID) { markPrescribed(ID, 2, 4); }
    void SetFamilyPrescribedQuaternion(unsigned int ID) { markPrescribed(ID, 3, 7); }
    void SetFamilyPrescribedQuaternion(unsigned int ID, const std::string& q_formula, bool dictate = true);
    void AddFamilyPrescribedAcc(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z);
    void AddFamilyPrescribedAngAcc(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z);
    void ChangeFamily(unsigned int ID_from, unsigned int ID_to);
    /// Change the family of the clumps whose centre lies in a box region (API.h:1030-1043 of the reference); only
    /// clumps currently in one of `orig_fam` are touched when that set is not empty. Returns the number changed.
    size_t ChangeClumpFamily(unsigned int fam_num, const std::pair<double, double>& X = std::pair<double, double>(-1e30, 1e30),
                             const std::pair<double, double>& Y = std::pair<double, double>(-1e30, 1e30),
                             const std::pair<double, double>& Z = std::pair<double, double>(-1e30, 1e30),
                             const std::set<unsigned int>& orig_fam = std::set<unsigned int>());
    void ChangeFamilyWhen(unsigned int, unsigned int, const std::string&);
    /// Value of a prescription string at time t (what the integrator will be given); exposed for scripts and tests
    static double EvaluatePrescription(const std::string& expression, double t) { return TimeExpression(expression).Eval(t); }
    void SetFamilyExtraMargin(unsigned int N, float extra_size);
    /// "Corrections" are code strings added to the integrator (API.h:806-838 of the reference): they need run-time
    /// compilation like custom force models, so they throw.
    void CorrectFamilyLinVel(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z,
                             const std::string& pre = "none");
    void CorrectFamilyAngVel(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z,
                             const std::string& pre = "none");
    void CorrectFamilyPosition(unsigned int ID, const std::string& X, const std::string& Y, const std::string& Z,
                               const std::string& pre = "none");
    void CorrectFamilyQuaternion(unsigned int ID, const std::string& q_formula);

    // ---- contact wildcards = the history words of the force model (API.h:841-870 of the reference): for the frictional
    // model delta_tan_x, delta_tan_y, delta_tan_z, delta_time; none for the frictionless one.  The setters rewrite the
    // named word of every listed contact whose two owners' families match.
    void SetContactWildcardValue(const std::string& name, float val);
    void SetFamilyContactWildcardValueEither(unsigned int N, const std::string& name, float val);
    void SetFamilyContactWildcardValueBoth(unsigned int N, const std::string& name, float val);
    void SetFamilyContactWildcardValue(unsigned int N1, unsigned int N2, const std::string& name, float val);
    /// Declaring wildcards is how a custom force model names its own storage: refused for the built-in models unless the
    /// set is exactly the one the model already has
    void SetContactWildcards(const std::set<std::string>& wildcards);
    void SetOwnerWildcards(const std::set<std::string>& wildcards);
    void SetGeometryWildcards(const std::set<std::string>& wildcards);
    // owner / geometry wildcard access: no built-in model declares any, so every name is unknown and these throw, as the
    // reference does for an unknown name (API.h:936-1014, APIPublic.cpp:1042-1180)
    void SetOwnerWildcardValue(bodyID_t ownerID, const std::string& name, const std::vector<float>& vals);
    void SetOwnerWildcardValue(bodyID_t ownerID, const std::string& name, float val, size_t n = 1) {
        SetOwnerWildcardValue(ownerID, name, std::vector<float>(n, val));
    }
    void SetFamilyOwnerWildcardValue(unsigned int N, const std::string& name, const std::vector<float>& vals);
    void SetFamilyOwnerWildcardValue(unsigned int N, const std::string& name, float val) {
        SetFamilyOwnerWildcardValue(N, name, std::vector<float>(1, val));
    }
    void SetTriWildcardValue(bodyID_t geoID, const std::string& name, const std::vector<float>& vals);
    void SetSphereWildcardValue(bodyID_t geoID, const std::string& name, const std::vector<float>& vals);
    void SetAnalWildcardValue(bodyID_t geoID, const std::string& name, const std::vector<float>& vals);
    std::vector<float> GetOwnerWildcardValue(bodyID_t ownerID, const std::string& name, bodyID_t n = 1);
    std::vector<float> GetAllOwnerWildcardValue(const std::string& name);
    std::vector<float> GetFamilyOwnerWildcardValue(unsigned int N, const std::string& name);
    std::vector<float> GetTriWildcardValue(bodyID_t geoID, const std::string& name, size_t n);
    std::vector<float> GetSphereWildcardValue(bodyID_t geoID, const std::string& name, size_t n);
    std::vector<float> GetAnalWildcardValue(bodyID_t geoID, const std::string& name, size_t n);
    void EnableOwnerWildcardOutput(bool enable = true) { (void)enable; }
    void EnableContactWildcardOutput(bool enable = true) {
        if (enable) m_cnt_out_content |= CNT_WILDCARD; else m_cnt_out_content &= ~(unsigned int)CNT_WILDCARD;
    }
    void EnableGeometryWildcardOutput(bool enable = true) { (void)enable; }

    // ---- persistent contacts (API.h:872-905 of the reference): a marked pair stays in the contact list even when the
    // broad phase no longer proposes it.  A built-in force model gives a pair that is not in touch no force and clears
    // its history (DEMCalcForceKernels.cu:258-262), so persistence changes no trajectory; it changes which POTENTIAL pairs
    // are reported.  The marks are kept by the facade (owner pairs) and honoured in GetContacts / GetClumpContacts /
    // GetContactDetailedInfo / WriteContactFileIncludingPotentialPairs, where a marked pair the list dropped re-appears
    // with zero force.  Like the reference they need a model with history.
    void MarkFamilyPersistentContactEither(unsigned int N);
    void MarkFamilyPersistentContactBoth(unsigned int N);
    void MarkFamilyPersistentContact(unsigned int N1, unsigned int N2);
    void MarkPersistentContact();
    void RemoveFamilyPersistentContactEither(unsigned int N);
    void RemoveFamilyPersistentContactBoth(unsigned int N);
    void RemoveFamilyPersistentContact(unsigned int N1, unsigned int N2);
    void RemovePersistentContact();
    size_t GetNumPersistentContacts() const { return m_persistent.size(); }

    void Initialize(bool dry_run = true);
    void DoDynamics(double thisCallDuration);
    void DoDynamicsThenSync(double thisCallDuration);
    void DoStepDynamics() { DoDynamics(m_ts_size); }
    void UpdateClumps();
    void ClearCache();

    void ShowThreadCollaborationStats();
    void ShowTimingStats();
    void ShowMemStats() const;
    void ShowAnomalies() {}
    void ClearThreadCollaborationStats() {}
    void ClearTimingStats() {}
    /// host bytes held by the facade's per-owner / per-geometry tables (the reference reports its two worker threads)
    size_t GetHostMemUsageDynamic() const;
    size_t GetHostMemUsageKinematic() const { return 0; }
    void PrintKinematicScratchSpaceUsage() const {}
    /// the reference waits for its asynchronous uploads here; every facade call returns with its transfer done
    void SyncMemoryTransfer() {}
    /// the reference re-derives bin size / margin policy and re-compiles; here changed settings take effect at once
    void UpdateSimParams();
    void ReleaseFlattenedArrays() {}
    /// empty in the reference too (APIPublic.cpp:2444)
    void PurgeFamily(unsigned int) {}
    /// "not implemented and has no effect" in the reference (APIPublic.cpp:803-820): the same here, with its warning
    void SetAdaptiveTimeStepType(const std::string& type);
    bool GetWhetherForceCollectInKernel() const { return true; }
    // run-time compilation settings of the reference: kept so that scripts can round-trip them; nothing is compiled here
    std::unordered_map<std::string, std::string> GetJitStringSubs() const { return {}; }
    std::vector<std::string> GetJitifyOptions() const { return m_jitify_options; }
    void SetJitifyOptions(const std::vector<std::string>& options) { m_jitify_options = options; }
    void SetKernelInclude(const std::string& includes) { m_kernel_includes = includes; }
    void AddKernelInclude(const std::string& lib_name) { m_kernel_includes += "#include <" + lib_name + ">\n"; }
    void RemoveKernelInclude() { m_kernel_includes = " "; }

    void WriteSphereFile(const std::filesystem::path& outfilename) const;
    void WriteClumpFile(const std::filesystem::path& outfilename, unsigned int accuracy = 10) const;
    void WriteContactFile(const std::filesystem::path& outfilename, float force_thres = 1e-15) const;
    void WriteMeshFile(const std::filesystem::path& outfilename) const;
    void WriteContactFileIncludingPotentialPairs(const std::filesystem::path& outfilename) const { WriteContactFile(outfilename, -1.0f); }
    void SetContactOutputFormat(OUTPUT_FORMAT format);
    void SetContactOutputFormat(const std::string& format);
    /// Clumps of this family are left out of the clump / sphere files
    void DisableFamilyOutput(unsigned int ID) { m_no_output_families.insert((family_t)ID); }
    static std::unordered_map<std::string, std::vector<float3>> ReadClumpFloat3FromCsv(
        const std::string& infilename, const std::string& x_header, const std::string& y_header, const std::string& z_header,
        const std::string& clump_header = "clump_type") {
        return ReadClumpXyzFromCsv(infilename, clump_header, x_header, y_header, z_header);
    }  // legacy-ASCII VTK of all meshes, current pose
    static std::unordered_map<std::string, std::vector<float3>> ReadClumpXyzFromCsv(
        const std::string& infilename, const std::string& clump_header = "clump_type", const std::string& x_header = "X",
        const std::string& y_header = "Y", const std::string& z_header = "Z");
    static std::unordered_map<std::string, std::vector<float4>> ReadClumpQuatFromCsv(
        const std::string& infilename, const std::string& clump_header = "clump_type", const std::string& qw_header = "Qw",
        const std::string& qx_header = "Qx", const std::string& qy_header = "Qy", const std::string& qz_header = "Qz");

    static std::unordered_map<std::string, std::vector<float3>> ReadClumpVelFromCsv(const std::string& infilename) {
        return ReadClumpXyzFromCsv(infilename, "clump_type", "v_x", "v_y", "v_z");
    }
    static std::unordered_map<std::string, std::vector<float3>> ReadClumpAngVelFromCsv(const std::string& infilename) {
        return ReadClumpXyzFromCsv(infilename, "clump_type", "w_x", "w_y", "w_z");
    }
    /// All contact pairs (geometry ids) of one contact type from a contact file (API.h:1190-1211 of the reference)
    static std::vector<std::pair<bodyID_t, bodyID_t>> ReadContactPairsFromCsv(
        const std::string& infilename, const std::string& cntType = "SS", const std::string& cntColName = "contact_type",
        const std::string& first_name = "geoA", const std::string& second_name = "geoB");
    /// All contact wildcards (every column that is not a standard contact-file column) of one contact type
    static std::unordered_map<std::string, std::vector<float>> ReadContactWildcardsFromCsv(
        const std::string& infilename, const std::string& cntType = "SS", const std::string& cntColName = "contact_type");

    // ---- deforming meshes (API.h:489-498 of the reference): new / incremented node positions in the mesh frame
    void SetTriNodeRelPos(size_t owner, size_t triID, const std::vector<float3>& new_nodes);
    void UpdateTriNodeRelPos(size_t owner, size_t triID, const std::vector<float3>& updates);
    std::vector<float3> GetMeshNodesGlobal(bodyID_t ownerID);
    std::shared_ptr<DEMMeshConnected> GetCachedMesh(bodyID_t ownerID);

    // ---- contact queries (API.h:500-570, 912-940 of the reference). GetContacts-like methods report POTENTIAL contacts
    // (every listed pair, like WriteContactFileIncludingPotentialPairs); owner-id pairs sorted by the A owner.
    std::vector<std::pair<bodyID_t, bodyID_t>> GetContacts() const;
    std::vector<std::pair<bodyID_t, bodyID_t>> GetContacts(const std::set<family_t>& family_to_include) const;
    std::vector<std::pair<bodyID_t, bodyID_t>> GetClumpContacts() const;
    std::vector<std::pair<bodyID_t, bodyID_t>> GetClumpContacts(const std::set<family_t>& family_to_include) const;
    /// ... also reporting the two owners' families
    std::vector<std::pair<bodyID_t, bodyID_t>> GetContacts(std::vector<std::pair<family_t, family_t>>& family_pair) const;
    std::vector<std::pair<bodyID_t, bodyID_t>> GetClumpContacts(std::vector<std::pair<family_t, family_t>>& family_pair) const;
    /// Every listed contact whose force is at least force_thres (negative: all potential pairs), with the fields
    /// SetContactOutputContent selected (API.h:552-569, dT.cpp:1619-1755 of the reference)
    std::shared_ptr<ContactInfoContainer> GetContactDetailedInfo(float force_thres = -1.0) const;
    /// Every force pair (contact point in the world frame, force on the queried owner) that concerns one of the owners;
    /// a contact between two listed owners is reported once, for the A side. Needs the force record (default on).
    size_t GetOwnerContactForces(const std::vector<bodyID_t>& ownerIDs, std::vector<float3>& points,
                                 std::vector<float3>& forces) const;
    size_t GetOwnerContactForces(const std::vector<bodyID_t>& ownerIDs, std::vector<float3>& points,
                                 std::vector<float3>& forces, std::vector<float3>& torques,
                                 bool torque_in_local = false) const;

    // raw owner access used by trackers (src/DEM/dT.cpp:3062-3130)
    /// state of n consecutive owners starting at ownerID (src/DEM/API.h:431-458 of the reference)
    std::vector<float3> GetOwnerPosition(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float3> GetOwnerVelocity(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float3> GetOwnerAngVel(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float4> GetOwnerOriQ(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float3> GetOwnerAcc(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float3> GetOwnerAngAcc(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<unsigned int> GetOwnerFamily(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float> GetOwnerMass(bodyID_t ownerID, bodyID_t n = 1) const;
    std::vector<float3> GetOwnerMOI(bodyID_t ownerID, bodyID_t n = 1) const;
    /// set the state of consecutive owners starting at ownerID, one per element (src/DEM/API.h:460-471)
    void SetOwnerPosition(bodyID_t ownerID, const std::vector<float3>& pos);
    void SetOwnerVelocity(bodyID_t ownerID, const std::vector<float3>& vel);
    void SetOwnerAngVel(bodyID_t ownerID, const std::vector<float3>& angVel);
    void SetOwnerOriQ(bodyID_t ownerID, const std::vector<float4>& oriQ);
    void SetOwnerFamily(bodyID_t ownerID, unsigned int fam, bodyID_t n = 1);
    /// extra (angular, owner frame) acceleration of consecutive owners for the next step only (API.h:477-486)
    void AddOwnerNextStepAcc(bodyID_t ownerID, const std::vector<float3>& acc);
    void AddOwnerNextStepAngAcc(bodyID_t ownerID, const std::vector<float3>& angAcc);
    void ChangeClumpSizes(const std::vector<bodyID_t>& IDs, const std::vector<float>& factors);
    double Reduce(int kind) const;
    double ReduceInRegion(int kind, const ScalarExpression& region) const;
    DemCtx* GetCoreContext() const { return ctx; }

  private:
    struct Prescription {
        bool used = false;
        bool linVelP[3] = {false, false, false}, rotVelP[3] = {false, false, false}, linPosP[3] = {false, false, false};
        bool rotPosP = false;
        bool hasLinVel[3] = {false, false, false}, hasRotVel[3] = {false, false, false}, hasLinPos[3] = {false, false, false};
        bool hasAcc[3] = {false, false, false}, hasAngAcc[3] = {false, false, false};
        float linVel[3] = {0, 0, 0}, rotVel[3] = {0, 0, 0}, linPos[3] = {0, 0, 0}, acc[3] = {0, 0, 0}, angAcc[3] = {0, 0, 0};
        // time-dependent components (nullptr = constant): re-evaluated before every step (DEM/utils/Expression.hpp)
        std::shared_ptr<TimeExpression> eLinVel[3], eRotVel[3], eLinPos[3], eAcc[3], eAngAcc[3];
        bool TimeDependent() const {
            for (int k = 0; k < 3; k++)
                if (eLinVel[k] || eRotVel[k] || eLinPos[k] || eAcc[k] || eAngAcc[k]) return true;
            return false;
        }
    };
    bool anyTimeDependentPrescription() const;
    void markPrescribed(unsigned int ID, int what, int axes);
    std::set<family_t> m_no_output_families;
    double simTimeOrZero() const;
    void check(int rc, const char* what) const;
    void uploadFamilies();
    void assertInit(const char* what) const;

    DemCtx* ctx = nullptr;
    VERBOSITY verbosity = INFO;
    unsigned int m_out_content = QUAT | ABSV;
    OUTPUT_FORMAT m_out_format = OUTPUT_FORMAT::CSV, m_cnt_out_format = OUTPUT_FORMAT::CSV;
    MESH_FORMAT m_mesh_out_format = MESH_FORMAT::VTK;
    void warnIfBinary(OUTPUT_FORMAT f, const char* what) const;
    unsigned int m_cnt_out_content = OWNER | FORCE | CNT_POINT;
    float3 G = make_float3(0, 0, -9.81f);
    double m_ts_size = 1e-5;
    int m_cd_update_freq = 20;
    bool m_adaptive_update_freq = true;
    unsigned int m_max_update_freq = 200;
    TIME_INTEGRATOR m_integrator = TIME_INTEGRATOR::EXTENDED_TAYLOR;
    FORCE_MODEL m_force_model = FORCE_MODEL::HERTZIAN;
    std::shared_ptr<DEMForceModel> m_force_model_obj = std::make_shared<DEMForceModel>(FORCE_MODEL::HERTZIAN);
    std::vector<std::string> m_jitify_options;
    std::string m_kernel_includes;
    // persistent contacts: (owner A, owner B, geometry A, geometry B, type) of every marked pair
    struct PersistentPair {
        bodyID_t ownerA, ownerB, geoA, geoB;
        uint8_t type;
        bool operator<(const PersistentPair& o) const {
            return std::tie(geoA, geoB, type) < std::tie(o.geoA, o.geoB, o.type);
        }
    };
    std::set<PersistentPair> m_persistent;
    std::vector<std::tuple<uint32_t, uint32_t, uint8_t>> persistentKeys() const;
    std::shared_ptr<ContactInfoContainer> generateContactInfo(float force_thres, unsigned int content) const;
    bool m_any_rolling_resistance = false;
    std::vector<unsigned char> m_sp_blob;  // the DemSimParams of Initialize(), for UpdateSimParams
    void markPersistent(int mode, unsigned int N1, unsigned int N2, bool mark);
    void setContactWildcard(int mode, unsigned int N1, unsigned int N2, const std::string& name, float val);
    [[noreturn]] void noSuchWildcard(const char* what, const std::string& name) const;
    float m_expand_factor = -1.f;
    float m_approx_max_vel = 1e15f;
    float m_expand_safety_multi = 1.f;
    float m_expand_base_vel = 3.f;
    float threshold_error_out_vel = 1e3f;
    float threshold_error_out_num_cnts = 100.f;
    bool no_recording_contact_forces = false;
    bool sys_initialized = false;
    float3 m_user_box_min = make_float3(-10, -10, -10), m_user_box_max = make_float3(10, 10, 10);
    float3 m_target_box_min = make_float3(-12, -12, -12), m_target_box_max = make_float3(12, 12, 12);
    int m_box_dir_exact = -1;  // axis along which the world spans the user's box exactly (-1: none)
    std::string m_user_add_bounding_box = "none";
    std::shared_ptr<DEMMaterial> m_bounding_box_material;

    std::vector<std::shared_ptr<DEMMaterial>> m_loaded_materials;
    std::map<std::string, std::map<std::pair<unsigned, unsigned>, float>> m_pairwise_matprop;
    std::vector<std::shared_ptr<DEMClumpTemplate>> m_templates;
    std::vector<std::shared_ptr<DEMClumpBatch>> m_cached_input_clump_batches;
    std::vector<std::shared_ptr<DEMExternObj>> m_cached_extern_objs;
    std::vector<std::shared_ptr<DEMMeshConnected>> m_cached_meshes;
    std::vector<std::shared_ptr<DEMTracker>> m_trackers;
    std::vector<std::pair<unsigned, unsigned>> m_input_no_contact_pairs;
    std::map<unsigned, Prescription> m_prescriptions;
    std::map<unsigned, float> m_family_extra_margin;
    std::vector<uint8_t> m_family_masks;

    size_t nOwnerClumps = 0, nOwnerBodies = 0, nSpheres = 0;
    // state of the already initialised owners carried over by UpdateClumps (empty otherwise)
    struct CarriedState {
        size_t nClumps = 0, nBodies = 0, nBatches = 0;
        std::vector<uint64_t> voxel;
        std::vector<uint16_t> lx, ly, lz;
        std::vector<float> quat, vel, omg;
        std::vector<uint8_t> fam;
        std::vector<uint32_t> idA, idB;
        std::vector<uint8_t> ctype;
        std::vector<float> wildcards;
    };
    CarriedState m_carry;
    size_t m_n_init_batches = 0;
    std::vector<float> m_owner_mass;
    std::vector<float3> m_owner_moi;
    std::vector<unsigned int> m_owner_type_mark;  // clump template mark per clump owner
    std::vector<unsigned int> m_sphere_owner, m_tri_owner, m_anal_owner;
    std::vector<unsigned int> m_owner_first_sphere;  // id of the first sphere of every owner (nSpheres for owners without)
    bodyID_t geoOwner(uint32_t geo, uint8_t type, bool sideB) const;
    std::vector<std::pair<bodyID_t, bodyID_t>> contactOwnerPairs(bool clumps_only, const std::set<family_t>* fams) const;
    double m_wall_time_dynamics = 0.0;
};

}  // namespace deme
