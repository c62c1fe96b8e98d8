// Host-only checks of facade pieces that never touch the device: region expressions of inspectors, the per-contact read-out
// container, the vector-form setters of a clump batch, the force-model handle, the small frame / quaternion helpers.
// Prints one "ok <name>" line per passed group; any failed expectation aborts with a message and a non-zero exit code.
#include <DEM/API.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>

using namespace deme;

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

int main() {
    {  // region strings in the form the reference's demos write them (DEMdemo_Repose.cpp, DEMdemo_Centrifuge.cpp)
        ScalarExpression r("return (abs(X) <= 0.48) && (abs(Y) <= 0.48) && (Z <= -0.44);", {"X", "Y", "Z"});
        const double in[3] = {0.1, -0.47, -0.5}, out1[3] = {0.49, 0, -0.5}, out2[3] = {0, 0, -0.43};
        EXPECT(r.Eval(in) != 0.0 && r.Eval(out1) == 0.0 && r.Eval(out2) == 0.0 && !r.IsConstant());
        ScalarExpression c("return (X * X + Y * Y <= 0.25 * 0.25) && Z > 0 ? 1 : 0;", {"X", "Y", "Z"});
        const double a[3] = {0.1, 0.2, 0.3}, b[3] = {0.2, 0.2, 0.3};
        EXPECT(c.Eval(a) == 1.0 && c.Eval(b) == 0.0);
        EXPECT(throws([] { ScalarExpression("return W > 0;", {"X", "Y", "Z"}); }, "W"));
        // an inspector whose region does not depend on the position is refused (AuxClasses.cpp:209-221 of the reference)
        EXPECT(throws([] { DEMInspector(nullptr, "clump_max_z", "return 1 > 0;"); }, "X, Y or Z"));
        EXPECT(throws([] { DEMInspector(nullptr, "no_such_quantity"); }, "not a known query type"));
        DEMInspector ok(nullptr, "clump_volume", "return Z < 0;");
        (void)ok;
        puts("ok regions");
    }
    {  // ContactInfoContainer: only the fields that are switched on exist (Structs.h:1049-1107)
        ContactInfoContainer info(OWNER | FORCE | CNT_WILDCARD, {"delta_tan_x", "delta_time"});
        info.ResizeAll(3);
        EXPECT(info.Size() == 3 && info.GetForce().size() == 3 && info.GetAOwner().size() == 3 && info.GetBOwner().size() == 3);
        EXPECT(info.GetAOwnerFamily().size() == 3 && info.GetContactType().size() == 3);
        EXPECT(info.GetWildcard("delta_time").size() == 3);
        info.GetWildcard("delta_time")[2] = 0.25f;
        EXPECT(info.GetWildcard("delta_time")[2] == 0.25f && info.GetWildcard("delta_tan_x")[2] == 0.f);
        EXPECT(throws([&] { info.GetPoint(); }, "does not have field: 'Point'"));
        EXPECT(throws([&] { info.GetNormal(); }, "SetContactOutputContent"));
        EXPECT(throws([&] { info.GetAGeo(); }) && throws([&] { info.GetTorque(); }));
        EXPECT(throws([&] { info.GetWildcard("delta_tan_y"); }, "delta_tan_y"));
        EXPECT(info.Contains(FORCE) && !info.Contains(NORMAL));
        info.ResizeAll(1);
        EXPECT(info.Size() == 1 && info.GetWildcard("delta_tan_x").size() == 1);
        ContactInfoContainer none(CNT_POINT, {"delta_time"});
        EXPECT(throws([&] { none.GetWildcard("delta_time"); }));
        puts("ok contact_info");
    }
    {  // clump batch: {x, y, z} for all and {{x, y, z}, ...} forms, wildcard bookkeeping (Structs.h:763-917)
        DEMClumpBatch b(3);
        b.SetPos(std::vector<float>{1, 2, 3});
        EXPECT(b.xyz.size() == 3 && b.xyz[2].y == 2.f);
        b.SetVel(std::vector<std::vector<float>>{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}});
        EXPECT(b.vel[1].y == 1.f && b.vel[2].z == 1.f);
        b.SetOriQ(std::vector<float>{0, 0, 0, 1});
        EXPECT(b.oriQ[0].w == 1.f);
        EXPECT(throws([&] { b.SetPos(std::vector<float>{1, 2}); }, "3-element"));
        EXPECT(throws([&] { b.SetVel(std::vector<std::vector<float>>{{1, 0, 0}}); }, "length 3"));
        EXPECT(throws([&] { b.SetFamilies(std::vector<unsigned int>{0, 1}); }));
        b.SetExistingContacts({{0, 1}, {1, 2}});
        b.AddExistingContactWildcard("delta_time", {0.1f, 0.2f});
        EXPECT(b.GetNumContacts() == 2 && b.contact_wildcards.at("delta_time")[1] == 0.2f);
        EXPECT(throws([&] { b.AddExistingContactWildcard("delta_tan_x", {0.1f}); }, "one value per existing contact"));
        b.AddOwnerWildcard("gran_strain", 0.5f);
        EXPECT(b.owner_wildcards.at("gran_strain").size() == 3);
        EXPECT(throws([&] { b.AddOwnerWildcard("gran_strain", std::vector<float>{1.f}); }));
        puts("ok clump_batch");
    }
    {  // the force-model handle: the built-in models keep their own history words, custom source is refused with a reason
        DEMForceModel m(FORCE_MODEL::HERTZIAN);
        m.SetPerContactWildcards({"delta_time", "delta_tan_x", "delta_tan_y", "delta_tan_z"});
        EXPECT(throws([&] { m.SetPerContactWildcards({"delta_time", "my_own"}); }, "run-time compilation"));
        EXPECT(throws([&] { m.SetPerOwnerWildcards({"gran_strain"}); }, "run-time compilation"));
        m.SetPerOwnerWildcards({});
        m.SetPerGeometryWildcards({});
        EXPECT(throws([&] { m.DefineCustomModel("force = 0;"); }, "run-time compilation"));
        EXPECT(throws([&] { m.SetForceModelType(FORCE_MODEL::CUSTOM); }));
        m.SetForceModelType(FORCE_MODEL::HERTZIAN_FRICTIONLESS);
        EXPECT(throws([&] { m.SetPerContactWildcards({"delta_time"}); }));
        m.SetPerContactWildcards({});
        m.SetMustHaveMatProp({"E", "nu"});
        EXPECT(m.m_must_have_mat_props.count("nu") == 1);
        puts("ok force_model");
    }
    {  // quaternion / frame helpers (src/DEM/HostSideHelpers.hpp:321-361, 585-633 of the reference)
        const float3 z = make_float3(0, 0, 1);
        const float4 q90 = QuatFromAxisAngle(z, (float)(PI / 2));
        const float3 r = Rotate(make_float3(1, 0, 0), q90);
        EXPECT(close(r.x, 0, 1e-6) && close(r.y, 1) && close(r.z, 0, 1e-6));
        const float4 q180 = hostHamiltonProduct(q90, q90);
        const float3 r2 = Rotate(make_float3(1, 0, 0), q180);
        EXPECT(close(r2.x, -1) && close(r2.y, 0, 1e-6));
        const float4 turned = RotateQuat(q90, z, (float)(PI / 2));  // a further quarter turn about the global z
        EXPECT(close(turned.z, q180.z) && close(turned.w, q180.w, 1e-6));
        const float3 rod = Rodrigues(make_float3(1, 0, 0), z, (float)(PI / 2));
        EXPECT(close(rod.x, 0, 1e-6) && close(rod.y, 1));
        const float3 rod2 = Rodrigues(make_float3(0.3f, -0.2f, 0.9f), make_float3(0, 1, 0), 0.7f);
        const float3 viaq = Rotate(make_float3(0.3f, -0.2f, 0.9f), QuatFromAxisAngle(make_float3(0, 1, 0), 0.7f));
        EXPECT(close(rod2.x, viaq.x, 1e-5) && close(rod2.y, viaq.y, 1e-5) && close(rod2.z, viaq.z, 1e-5));
        const std::vector<double> p = {0.3, -1.2, 2.0}, shift = {1, 2, 3};
        const std::vector<double> qd = {q90.x, q90.y, q90.z, q90.w};
        const std::vector<double> g = FrameTransformLocalToGlobal(p, shift, qd);
        EXPECT(close(g[0], 1.2 + 1, 1e-6) && close(g[1], 0.3 + 2, 1e-6) && close(g[2], 5.0, 1e-6));
        const std::vector<double> back = FrameTransformGlobalToLocal(g, shift, qd);
        EXPECT(close(back[0], p[0], 1e-6) && close(back[1], p[1], 1e-6) && close(back[2], p[2], 1e-6));
        EXPECT(sign_func(-2.5) == -1 && sign_func(0) == 0 && sign_func(3u) == 1);
        EXPECT(vector_sum(std::vector<int>{1, 2, 3}) == 6);
        EXPECT(isBetween(make_float3(0, 0, 0), make_float3(-1, -1, -1), make_float3(1, 1, 1)));
        EXPECT(!isBetween(make_float3(0, 2, 0), make_float3(-1, -1, -1), make_float3(1, 1, 1)));
        EXPECT(str_to_upper("xYz_1") == "XYZ_1");
        EXPECT((hostRemoveElem(std::vector<int>{1, 2, 3, 4}, std::vector<bool>{false, true, false, true}) == std::vector<int>{1, 3}));
        EXPECT((parse_string_line("a,b,,c") == std::vector<std::string>{"a", "b", "", "c"}));
        EXPECT(check_exist(std::set<int>{1, 2}, 2) && !check_exist(std::vector<int>{1, 2}, 3));
        EXPECT((Real4ToVec(make_float4(1, 2, 3, 4)) == std::vector<float>{1, 2, 3, 4}));
        puts("ok helpers");
    }
    {  // clump template frames, mesh accessors, Wavefront round trip, Merge
        DEMClumpTemplate t;
        t.relPos = {make_float3(1, 0, 0), make_float3(0, 2, 0), make_float3(0.5f, 0.5f, 3)};
        t.radii = {1, 1, 1};
        t.nComp = 3;
        const std::vector<float3> before = t.relPos;
        const float4 q = QuatFromAxisAngle(make_float3(0, 0, 1), 0.9f);
        t.Move(make_float3(0.3f, -0.2f, 0.1f), q);
        EXPECT(!close(t.relPos[0].x, before[0].x, 1e-3));
        t.InformCentroidPrincipal(std::vector<float>{0.3f, -0.2f, 0.1f}, std::vector<float>{q.x, q.y, q.z, q.w});  // the inverse
        for (size_t i = 0; i < 3; i++)
            EXPECT(close(t.relPos[i].x, before[i].x, 1e-6) && close(t.relPos[i].y, before[i].y, 1e-6) && close(t.relPos[i].z, before[i].z, 1e-6));
        EXPECT(throws([&] { t.Move(std::vector<float>{1, 2}, std::vector<float>{0, 0, 0, 1}); }, "3-element"));

        DEMMeshConnected a, b;
        a.SetGeometry({make_float3(0, 0, 0), make_float3(1, 0, 0), make_float3(0, 1, 0), make_float3(0, 0, 1)},
                      {make_int3(0, 1, 2), make_int3(0, 1, 3)});
        b.SetGeometry({make_float3(5, 0, 0), make_float3(6, 0, 0), make_float3(5, 1, 0)}, {make_int3(0, 1, 2)});
        b.m_normals = {make_float3(0, 0, 1)};
        b.m_face_n_indices = {make_int3(0, 0, 0)};
        EXPECT(a.GetNumTriangles() == 2 && a.GetTriangle(1).p3.z == 1.f && a.GetIndicesVertexesAsVectorOfVectors()[1][2] == 3);
        EXPECT(a.GetCoordsVerticesAsVectorOfVectors()[1][0] == 1.f && a.GetCoordsNormals().empty() && b.GetIndicesNormals().size() == 1);
        EXPECT(throws([&] { a.GetTriangle(2); }));
        a.AddGeometryWildcard("wear", 0.f);
        EXPECT(a.geo_wildcards.at("wear").size() == 2);
        EXPECT(throws([&] { a.AddGeometryWildcard("wear", std::vector<float>{1.f}); }, "2 triangles"));
        a.ClearWildcards();
        std::vector<DEMMeshConnected> both = {a, b};
        const char* path = "/tmp/facade_host_check_merged.obj";
        DEMMeshConnected::WriteWavefront(path, both);
        DEMMeshConnected back;
        EXPECT(back.LoadWavefrontMesh(path));
        EXPECT(back.GetNumNodes() == 7 && back.GetNumTriangles() == 3);
        EXPECT(back.GetIndicesVertexes()[2].x == 4 && back.GetIndicesVertexes()[2].z == 6 && back.GetCoordsVertices()[6].y == 1.f);
        DEMMeshConnected merged = DEMMeshConnected::Merge(both);
        EXPECT(merged.GetNumNodes() == 7 && merged.GetNumTriangles() == 3 && merged.GetIndicesVertexes()[2].y == 5);
        EXPECT(merged.GetCoordsNormals().empty());  // only one of the two had normals
        for (size_t i = 0; i < 3; i++) {
            const int3 f = merged.GetIndicesVertexes()[i], g = back.GetIndicesVertexes()[i];
            EXPECT(f.x == g.x && f.y == g.y && f.z == g.z);
        }
        auto mat = std::make_shared<DEMMaterial>(std::unordered_map<std::string, float>{{"E", 1e8f}});
        both[0].SetMaterial(mat);
        both[1].SetMaterial(mat);
        EXPECT(DEMMeshConnected::Merge(both).materials.size() == 3);
        EXPECT(throws([] { DEMInspector(nullptr, "absv").SetInspectionCode("quantity[myOwner] = 1;"); }, "run time"));
        puts("ok mesh_and_templates");
    }
    {  // input files as they come: Windows line endings, padded cells, a template file without a radius column
        FILE* f = fopen("/tmp/facade_host_check_crlf.csv", "wb");
        fputs("X,Y,Z, Qw,Qx,Qy,Qz,clump_type\r\n-0.6,-0,-0.01,1,0,0,0,0\r\n-0.58, 0.25 ,-0.01,1,0,0,0,1\r\n\r\n0.5,0,0,1,0,0,0,0\r\n", f);
        fclose(f);
        auto xyz = DEMSolver::ReadClumpXyzFromCsv("/tmp/facade_host_check_crlf.csv");
        EXPECT(xyz.size() == 2 && xyz.at("0").size() == 2 && xyz.at("1").size() == 1 && close(xyz.at("1")[0].y, 0.25) && close(xyz.at("0")[1].x, 0.5));
        auto quat = DEMSolver::ReadClumpQuatFromCsv("/tmp/facade_host_check_crlf.csv");
        EXPECT(quat.at("0").size() == 2 && close(quat.at("0")[0].w, 1));
        f = fopen("/tmp/facade_host_check_noradius.csv", "wb");
        fputs("x,y,z\r\n1,2,3\r\n4,5,6\r\n", f);
        fclose(f);
        DEMClumpTemplate t;
        t.ReadComponentFromFile("/tmp/facade_host_check_noradius.csv");  // (the reference reads with ignore_missing_column)
        EXPECT(t.nComp == 2 && t.radii.size() == 2 && t.radii[1] == 0.f && close(t.relPos[1].y, 5));
        t.ReadComponentFromFile("/tmp/facade_host_check_noradius.csv");  // ... and appends
        EXPECT(t.nComp == 4 && t.relPos.size() == 4 && close(t.relPos[2].x, 1));
        EXPECT(throws([&] { t.ReadComponentFromFile("/nonexistent/file.csv"); }, "cannot be opened"));
        puts("ok input_files");
    }
    return 0;
}
