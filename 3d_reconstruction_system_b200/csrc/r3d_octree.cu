// r3d_octree.cu -- occupancy entry points (placeholder bodies until the brick store lands).
#include "r3d_common.cuh"
using namespace r3d;
#define R3D_TODO(ctxexpr) return set_error(ctxexpr, R3D_ERR_UNSUPPORTED, "%s: not implemented yet", __func__)
struct r3d_tree { r3d_ctx* ctx; double res; };
extern "C" int r3d_tree_create(r3d_ctx* ctx, double resolution, r3d_tree** tree) { (void)resolution; (void)tree; R3D_TODO(ctx); }
extern "C" void r3d_tree_destroy(r3d_tree* tree) { delete tree; }
extern "C" int r3d_tree_clear(r3d_tree* t) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_params(r3d_tree* t, float*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_update_points(r3d_tree* t, const float*, uint64_t, int, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_update_points_f64(r3d_tree* t, const double*, uint64_t, int, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_update_points_logodds(r3d_tree* t, const float*, uint64_t, float, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_insert_scan(r3d_tree* t, const float*, uint64_t, const float*, double, int) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_scan_delta_compute(r3d_tree* t, const float*, uint64_t, const float*, double, int, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_scan_delta_export(r3d_tree* t, void*, uint64_t, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_apply_delta(r3d_tree* t, const void*, uint64_t) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_delta_expand_keys(const void*, uint64_t, uint16_t*, uint64_t, uint64_t*, uint16_t*, uint64_t, uint64_t*) { R3D_TODO(nullptr); }
extern "C" int r3d_tree_update_inner_occupancy(r3d_tree* t) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_write_bt(r3d_tree* t, const char*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_write_bt_mem(r3d_tree* t, uint8_t*, size_t, size_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_to_max_likelihood(r3d_tree* t) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_num_voxels(r3d_tree* t, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_size(r3d_tree* t, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_search(r3d_tree* t, const uint16_t*, uint64_t, float*, uint8_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_tree_export_voxels(r3d_tree* t, uint16_t*, float*, uint64_t, uint64_t*) { R3D_TODO(t ? t->ctx : nullptr); }
extern "C" int r3d_coord_to_key(r3d_tree* t, const float*, uint64_t, uint16_t*, uint8_t*) { R3D_TODO(t ? t->ctx : nullptr); }
