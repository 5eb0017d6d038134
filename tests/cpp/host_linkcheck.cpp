// Every function hanamaru_host.h declares must be defined in libhanamaru_host.so: built and linked by tests/test_host_link.py.
#include "hanamaru_host.h"
using namespace hanamaru;
int main() {
    volatile void* p[] = {
        (void*)&hsv_to_rgb, (void*)&parse_obj, (void*)&build_from_mesh, (void*)&build_from_scene,
        (void*)&init_scene_rtcamp6_v3_1, (void*)&init_scene_rtcamp6_v4, (void*)&init_scene_simple, (void*)&init_scene_material_examples,
        (void*)&init_scene_rtcamp6_v1, (void*)&init_scene_rtcamp6_v2, (void*)&init_scene_rtcamp6_v3, (void*)&init_scene_rtcamp5, (void*)&init_scene_tbf3,
        (void*)&init_scene_bvh_heavy, (void*)&init_scene_diamond, (void*)&init_scene_by_name, (void*)&scene_asset_paths, (void*)&core_api,
        (void*)&BvhScene::from_scene, (void*)&BvhMesh::from_mesh, (void*)&ObjLoader::load, (void*)&FlatSceneBuilder::flatten_tree,
    };
    (void)p;
    AssetStore a; std::string e; a.load_pack("x", &e);
    Matrix44 m = Matrix44::rotate_x(0.1) * Matrix44::rotate_y(0.2) * Matrix44::scale(1, 2, 3) * Matrix44::translate(1, 2, 3);
    Vector3 v = m * Vector3(1, 2, 3); (void)v;
    Camera c(Vector3(0, 0, 1), Vector3(0, 0, 0), Vector3(0, 1, 0), 20.0, LensShape::Circle, 0.1, 1.0); c.abi();
    StdRng r({1, 2}); r.gen_range(0, 1);
    PathTracingRenderer pr(1, 1.0, 1.0); DebugRenderer dr(2);
    FlatScene fs; fs.finalize(); FlatSceneBuilder fb(fs);
    Scene s; s.emissions();
    Aabb b{Vector3(0, 0, 0), Vector3(1, 1, 1)}; b.merge(b); b.intersect_aabb(b);
    return 0;
}
