// Host-only check of DEMMeshConnected (no GPU): Wavefront loading (triangles, quads, v/vt/vn corners, negative
// indices), Scale, Move, Mirror.  Prints vertices and faces for the Python test.
#include <DEM/API.h>
#include <DEM/HostSideHelpers.hpp>

#include <cstdio>

using namespace deme;

static void dump(const char* name, DEMMeshConnected& m) {
    printf("%s %zu %zu mass %.9g moi %.9g %.9g %.9g\n", name, m.GetNumNodes(), m.GetNumTriangles(), m.mass, m.MOI.x, m.MOI.y, m.MOI.z);
    for (const auto& v : m.GetCoordsVertices()) printf("v %.9g %.9g %.9g\n", v.x, v.y, v.z);
    for (const auto& f : m.GetIndicesVertexes()) printf("f %d %d %d\n", f.x, f.y, f.z);
}

int main(int argc, char** argv) {
    DEMMeshConnected mesh;
    if (!mesh.LoadWavefrontMesh(argv[1])) return 1;
    mesh.SetMass(2.f);
    mesh.SetMOI(make_float3(1.f, 2.f, 3.f));
    dump("loaded", mesh);
    DEMMeshConnected scaled = mesh;
    scaled.Scale(2.f);
    dump("scaled", scaled);
    DEMMeshConnected moved = mesh;
    moved.Move(make_float3(1.f, 2.f, 3.f), QuatFromAxisAngle(make_float3(0, 0, 1), 1.57079632679f));
    dump("moved", moved);
    DEMMeshConnected mirrored = mesh;
    mirrored.Mirror(make_float3(0, 0, 0), make_float3(1, 0, 0));
    dump("mirrored", mirrored);
    DEMMeshConnected missing;
    printf("missing %d\n", (int)missing.LoadWavefrontMesh("/nonexistent/file.obj"));
    return 0;
}
