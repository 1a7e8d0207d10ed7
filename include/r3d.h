/*
 * r3d.h -- C ABI of the B200-native mapping hot path (libr3d_b200.so).
 *
 * Drop-in boundary for rainfall1998/3D_reconstruction_system's depth -> world points -> OctoMap
 * path.  The reference is a set of Python scripts; the seam a maintainer binds is ctypes
 * (see INTEGRATION.md).  Every entry point cites the reference interface it replaces; paths are
 * relative to the reference checkout.  Plain C types only: no torch / numpy / C++ types cross it.
 *
 * Conventions
 *   - return 0 (R3D_OK) on success, a negative R3D_ERR_* otherwise; r3d_last_error() has the text.
 *     No C++ exception crosses the ABI.
 *   - data pointers may be HOST or DEVICE memory of the context's GPU (detected with
 *     cudaPointerGetAttributes).  Host pointers are staged through the GPU inside the call.
 *   - the caller owns every buffer it passes; the library owns r3d_ctx / r3d_tree and their device
 *     scratch, released by the *_destroy calls.
 *   - a context is single-threaded (the reference is synchronous, single-threaded).  Calls block
 *     until their results are visible unless r3d_set_blocking(ctx, 0) was called (device pointers
 *     only); then r3d_synchronize() fences.
 *   - there is NO CPU fallback: with no CUDA device r3d_create() fails.
 */
#ifndef R3D_H_
#define R3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r3d_ctx r3d_ctx;
typedef struct r3d_tree r3d_tree;

enum {
    R3D_OK = 0,
    R3D_ERR_ARG = -1,         /* bad argument */
    R3D_ERR_CUDA = -2,        /* CUDA runtime error, text in r3d_last_error */
    R3D_ERR_OOM = -3,         /* device / host allocation failed */
    R3D_ERR_STATE = -4,       /* call sequence error */
    R3D_ERR_IO = -5,          /* file could not be written */
    R3D_ERR_UNSUPPORTED = -6  /* valid request this build does not serve */
};

/* depth sample type (a1: what cv.imread hands to gentxtcord) */
enum { R3D_U8 = 0, R3D_U16 = 1, R3D_F32 = 2 };
/* a2 (depth) vs a3 (disparity, Z = fB / (raw * depth_scale), raw*scale <= 0 -> Z = 0) */
enum { R3D_MODE_DEPTH = 0, R3D_MODE_DISPARITY = 1 };
/* xyz record type written by the back-projection kernel */
enum { R3D_OUT_F32 = 0, R3D_OUT_F64 = 1 };

/* ------------------------------------------------------------------ context ---- */
const char *r3d_version(void);
int r3d_device_count(void);
/* One context per GPU.  Returns NULL on failure (r3d_last_error(NULL) has the reason). */
r3d_ctx *r3d_create(int device);
void r3d_destroy(r3d_ctx *ctx);
const char *r3d_last_error(r3d_ctx *ctx);
int r3d_set_blocking(r3d_ctx *ctx, int blocking);
int r3d_synchronize(r3d_ctx *ctx);
/* cudaStream_t all kernels of this context are launched on (for CUDA-event timing by the caller). */
void *r3d_stream(r3d_ctx *ctx);
/* Number of kernels launched by this context so far. */
uint64_t r3d_launch_count(r3d_ctx *ctx);
/* Device time (ms, CUDA events on the context stream) of the dominant kernel of the last call. */
float r3d_last_kernel_ms(r3d_ctx *ctx);
/* Pinned host memory helpers (so callers can hand the library page-locked buffers). */
void *r3d_host_alloc(size_t bytes);
void r3d_host_free(void *p);
/* Device memory of the context's GPU, for callers that keep intermediate results (e.g. the world points of a frame
 * batch between back-projection and insertPointCloud) on the GPU.  r3d_memcpy copies between any two of host / device
 * buffers on the context stream and blocks until done. */
void *r3d_device_alloc(r3d_ctx *ctx, size_t bytes);
void r3d_device_free(r3d_ctx *ctx, void *p);
int r3d_memcpy(r3d_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ------------------------------------------------------------------ a1: PNG ---- */
/*
 * Depth / disparity PNG decode for whole frame batches on a host thread pool (zlib inflate + un-filter + conversion
 * straight into the caller's frame stack, which may be pinned memory from r3d_host_alloc).  Replaces, for PNG files,
 *   R3D_PNG_GRAY8   cv.imread(path, IMREAD_GRAYSCALE)          transfer/camera_to_world.py:160  (uint8; 16-bit >> 8)
 *   R3D_PNG_CHANNEL cv.imread(path, IMREAD_UNCHANGED)[:, :, c] transfer/pixel_to_camera.py:133-134 (c = 1: green)
 *   R3D_PNG_RAW     IMREAD_UNCHANGED, first channel when there are several (full-precision 16-bit depth / disparity)
 * bit for bit (pinned against OpenCV 4.13 in tests/test_png_cpu.py).  Every file must be W x H; elem_bytes is 1 for
 * GRAY8 and the file's sample width (1 or 2) otherwise.  frame_stride_bytes 0 = packed.  n_threads 0 = one per core.
 * status (optional, n ints) receives R3D_OK / R3D_ERR_IO per file.  Needs no GPU.
 */
enum { R3D_PNG_GRAY8 = 0, R3D_PNG_CHANNEL = 1, R3D_PNG_RAW = 2 };
int r3d_png_info(const char *path, int *W, int *H, int *channels, int *bit_depth);
int r3d_png_decode_batch(const char *const *paths, int n, int mode, int channel, void *out, size_t frame_stride_bytes,
                         int elem_bytes, int W, int H, int n_threads, int *status);

/* ------------------------------------------------------------------ a8: text in */
/*
 * txt_read of the OctoMap scripts on a host thread pool: the points of an `x,y,z` text file (comma_mode != 0:
 * octomap/txt_transfer_octomap.py:16-28, also what get_pointdata re-reads, transfer/camera_to_world.py:92-98) or of an
 * ASCII PLY body (comma_mode == 0: octomap/ply_transfer_octomap.py:16-40 -- skip_lines = 8, whitespace separated, first
 * three columns, max_points = 5 400 001).  Numbers are converted like float() (correctly rounded, same grammar).  Empty and
 * blank lines are skipped (documented deviation: the reference's own PLY writer ends files with one and its reader raises on
 * it); any other line the reference raises on -- a field float() rejects in ANY column, fewer than three fields -- returns
 * R3D_ERR_ARG with the line number, like the reference's ValueError.  Lines after max_points are not looked at.  out: capacity x 3
 * doubles, or NULL for a size query; *n_points is always set.  Needs no GPU.
 */
int r3d_read_xyz_text(const char *path, int skip_lines, int comma_mode, uint64_t max_points, double *out, uint64_t capacity,
                      uint64_t *n_points, int n_threads);

/* ------------------------------------------------------------------ poses ------ */
/*
 * Replaces scipy_transfer(quat) = np.matrix(R.from_quat(quat).as_matrix()).I
 * (transfer/camera_to_world.py:53-55) for n frames at once.
 * poses: n x 7 doubles (qx,qy,qz,qw, tx,ty,tz) -- scipy scalar-last order, any norm > 0
 *        (camera_to_world.py:156-157 feeds file columns 4..7 / 1..3 as is).
 * t_scale: ICP scale correction folded into the pose table, t <- t_scale * t (1.0 = reference).
 * rt:    n x 12 doubles out: R^-1 row-major (9) then t (3).
 * Host computation (per frame, not per point).  R3D_ERR_ARG on a zero-norm quaternion
 * (scipy raises ValueError).
 */
int r3d_pose_to_rt(const double *poses, int n, double t_scale, double *rt);

/* ------------------------------------------------------------------ K1 --------- */
/*
 * Fused depth decode -> pinhole back-projection -> pose transform -> packed xyz records.
 * Replaces, per frame, gentxtcord (transfer/camera_to_world.py:67-83, transfer/pixel_to_camera.py:24-44)
 * + get_pointdata/point_camera (transfer/camera_to_world.py:57-59, 86-105) minus their text files:
 *   X = ((u - cx)/fx) * Z,  Y = ((v - cy)/fy) * Z,  p_world = R^-1 (p_cam - t)
 * evaluated in fp64, every product/sum rounded separately in the reference's order, then cast once.
 *
 * depth:   n_frames images, H rows of W samples, row pitch `pitch` bytes (0 = W*sizeof(sample)),
 *          frames contiguous (frame stride = H*pitch).
 * intr:    fx, fy, cx, cy.
 * rt:      n_frames x 12 doubles from r3d_pose_to_rt, or NULL for camera-frame output (gentxtcord only).
 * compact: 0 = every pixel emitted in row-major order like the reference (Z = 0 included);
 *          1 = pixels whose decoded depth/disparity is not finite-positive are dropped, order kept.
 * out_xyz: n_frames*W*H records of 3 x (float|double); the float form is the body of a
 *          binary_little_endian PLY with `property float x/y/z`.
 * out_counts: optional, n_frames uint64: records written per frame.
 */
int r3d_backproject_rt(r3d_ctx *ctx, const void *depth, int dtype, int W, int H, size_t pitch, int n_frames,
                       const double intr[4], const double *rt, int mode, double depth_scale, double fB,
                       int compact, int out_dtype, void *out_xyz, uint64_t *out_counts);
/* Same, from quaternion+translation poses (n_frames x 7, see r3d_pose_to_rt), float32 records. */
int r3d_backproject(r3d_ctx *ctx, const void *depth, int dtype, int W, int H, size_t pitch, int n_frames,
                    const double intr[4], const double *poses, int mode, double depth_scale, double fB,
                    int compact, float *out_xyz, uint64_t *out_counts);
/*
 * point_camera(p1, r_inverse, t) (transfer/camera_to_world.py:57-59) for n camera-frame points that already
 * exist as numbers (e.g. re-read from ./point/<name>.txt by get_pointdata, :92-99):
 * out = R^-1 (p - t), same operation order as r3d_backproject_rt.  rt: 12 doubles; xyz in/out: n x 3 doubles.
 */
int r3d_pose_apply_points(r3d_ctx *ctx, const double *xyz, uint64_t n, const double rt[12], double *out_xyz);
/*
 * 4x4 homogeneous transform of a point cloud: other_tools/transfer_T_icp.py:10-12,71-97
 * (point_camera(p, T) = T . [x y z 1]^T, first three rows).  xyz in / out: n x 3 doubles.
 */
int r3d_transform_points(r3d_ctx *ctx, const double *xyz, uint64_t n, const double T[16], double *out_xyz);

/* The PNG decoder's own zlib-stream inflate (RFC 1950 / 1951, output size known in advance), exposed for tests:
 * same bytes as zlib's uncompress() for every stream it accepts, an error for truncated / corrupted ones. */
int r3d_inflate(const void *src, size_t src_len, void *dst, size_t dst_len);

/* ------------------------------------------------------------------ K6: text --- */
/*
 * The vertex rows of genply / genply_RGB (transfer/camera_to_world.py:112-134, transfer/pixel_to_camera.py:98-124):
 *     "%.4f %.4f %.4f \n"            per point (trailing space), or, with rgb != NULL, genply_noRGB's
 *     "%.4f %.4f %.4f %d %d %d 0\n"  (transfer/pixel_to_camera.py:55-91),
 * byte for byte what Python's "%.4f" prints (exact value, round half to even; "inf", "-inf", "nan").
 * Coordinate i is x[i*stride], y[i*stride], z[i*stride] (stride in doubles: 1 for three arrays, 3 for an interleaved
 * n x 3 block); rgb is n x 3 bytes.  Buffers may be host or device memory (a device `out` must be 16-byte aligned).
 * *len always receives the text length; nothing is written when out is NULL or cap < *len (size query).
 * The header / trailer around the rows are a few constant lines written by the caller.
 */
int r3d_format_ply_rows(r3d_ctx *ctx, const double *x, const double *y, const double *z, size_t stride, uint64_t n,
                        const uint8_t *rgb, char *out, size_t cap, size_t *len);

/*
 * The lines of the x,y,z txt files: gentxtcord's `str(X) + ',' + str(Y) + ',' + str(Z) + '\n'`
 * (transfer/camera_to_world.py:79-81, transfer/pixel_to_camera.py:41-42) and get_pointdata's world lines (:103-104).
 * Every field is Python's str(float64): the shortest decimal that reads back to the same double, laid out by repr()'s
 * rules ("0.1", "1e-05", "1.2345678901234568e+16", "-0.0", "inf", "nan").  z_is_integer != 0 prints Z as an integer
 * ("157"), which is what gentxtcord does because Z is still the raw integer sample there.  Same buffer and size-query
 * conventions as r3d_format_ply_rows; a row is at most 75 bytes.
 */
int r3d_format_txt_rows(r3d_ctx *ctx, const double *x, const double *y, const double *z, size_t stride, uint64_t n,
                        int z_is_integer, char *out, size_t cap, size_t *len);

/* ------------------------------------------------------------------ occupancy -- */
/*
 * octomap.OcTree(resolution) (octomap/txt_transfer_octomap.py:33, octomap/ply_transfer_octomap.py:45):
 * depth 16, float32 log-odds, upstream default sensor model (hit 0.7, miss 0.4, clamp 0.1192 / 0.971,
 * occupancy threshold 0.5).
 */
int r3d_tree_create(r3d_ctx *ctx, double resolution, r3d_tree **tree);
void r3d_tree_destroy(r3d_tree *tree);
int r3d_tree_clear(r3d_tree *tree);
/* Capacity hint (like std::vector::reserve): room for n_bricks 8x8x8-voxel bricks (2 112 B each) without regrowth. */
int r3d_tree_reserve(r3d_tree *tree, uint64_t n_bricks);
/* out[5] = hit, miss, clamp_min, clamp_max, occupancy threshold (float32 log-odds). */
int r3d_tree_params(r3d_tree *tree, float out[5]);
int r3d_tree_resolution(r3d_tree *tree, double *resolution);   /* tree.getResolution() */

/*
 * tree.updateNode(point, True|False) for n points in call order
 * (octomap/txt_transfer_octomap.py:25, octomap/ply_transfer_octomap.py:33): coordToKeyChecked on the
 * float32 point, out-of-range points silently dropped, leaf log-odds += hit|miss, clamped.
 * n_dropped (optional) receives the number of out-of-range points.
 */
int r3d_tree_update_points(r3d_tree *tree, const float *xyz, uint64_t n, int occupied, uint64_t *n_dropped);
/* Same for float64 input, cast to float32 per coordinate exactly as the Python binding does. */
int r3d_tree_update_points_f64(r3d_tree *tree, const double *xyz, uint64_t n, int occupied, uint64_t *n_dropped);
/* updateNode(point, float log_odds_update): the binding's non-bool overload. */
int r3d_tree_update_points_logodds(r3d_tree *tree, const float *xyz, uint64_t n, float log_odds_update,
                                   uint64_t *n_dropped);

/*
 * tree.insertPointCloud(points, origin, maxrange, lazy_eval=False, discretize) (upstream binding; named by
 * BASELINE.json north_star, no call site in the reference): ray-cast free cells with the 3-D DDA of
 * computeRayKeys, endpoint occupied when within maxrange (maxrange < 0: unlimited), occupied wins,
 * every key updated once (miss for free, hit for occupied), clamped.
 */
int r3d_tree_insert_scan(r3d_tree *tree, const float *xyz, uint64_t n, const float origin[3], double maxrange,
                         int discretize);
/* n_scans consecutive insertPointCloud calls in one library call (no per-scan trip through the caller's language):
 * scan s has n_points[s] points stored back to back in xyz (host or device) and origin origins[3 s .. 3 s + 2] (host).
 * r3d_tree_last_scan_stats then reports the totals of the batch. */
int r3d_tree_insert_scans(r3d_tree *tree, const float *xyz, const uint64_t *n_points, const float *origins, uint32_t n_scans,
                          double maxrange, int discretize);
/*
 * The two halves of r3d_tree_insert_scan, for multi-GPU merging:
 * compute the scan's delta (set of free / occupied voxels) without touching the tree, export it as
 * brick records, apply records (possibly received from another GPU) in scan order.
 * Record layout (R3D_DELTA_RECORD_BYTES = 136): uint64 brick key (bx | by<<13 | bz<<26, b = key>>3),
 * 16 x uint32 occupied mask, 16 x uint32 free mask (already minus occupied); bit index inside a brick =
 * Morton code of (kx&7, ky&7, kz&7) with x lowest.
 */
#define R3D_DELTA_RECORD_BYTES 136
int r3d_scan_delta_compute(r3d_tree *tree, const float *xyz, uint64_t n, const float origin[3], double maxrange,
                           int discretize, uint64_t *n_records);
/* exports the delta of the last r3d_scan_delta_compute / r3d_tree_insert_scan on this tree (r3d_tree_insert_scans applies
 * its records straight from the pipeline's buffers and leaves none: *n_records = 0) */
int r3d_scan_delta_export(r3d_tree *tree, void *records, uint64_t capacity_records, uint64_t *n_records);
int r3d_tree_apply_delta(r3d_tree *tree, const void *records, uint64_t n_records);
/* Multi-GPU apply with the map partitioned by brick: only the records whose brick is owned by `part` of `nparts`
 * (owner = (hash64(brick key) >> 32) mod nparts, hash64 = the murmur3 finaliser) are applied; the others are skipped. */
int r3d_tree_apply_delta_owned(r3d_tree *tree, const void *records, uint64_t n_records, uint32_t part, uint32_t nparts);
/*
 * Whole-brick transfer between maps (merging the per-GPU pieces of a partitioned map, checkpointing):
 * record layout (R3D_BRICK_RECORD_BYTES = 2120): uint64 brick key, 512 float32 log-odds (Morton order inside the
 * brick, as in the delta masks), 16 x uint32 "known" mask.  Import creates missing bricks; voxels known in the record
 * replace the local value.  Buffers may be host or device memory.
 */
#define R3D_BRICK_RECORD_BYTES 2120
int r3d_tree_num_bricks(r3d_tree *tree, uint64_t *n);
int r3d_tree_export_bricks(r3d_tree *tree, void *records, uint64_t capacity_records, uint64_t *n_records);
int r3d_tree_import_bricks(r3d_tree *tree, const void *records, uint64_t n_records);
/* Batched forms for the multi-GPU rounds: n_scans deltas computed back to back into one record buffer (counts[s] records
 * each; R3D_ERR_OOM when the buffer is too small), and n_scans deltas (device memory, back to back) applied in order. */
int r3d_scan_deltas_compute(r3d_tree *tree, const float *xyz, const uint64_t *n_points, const float *origins, uint32_t n_scans,
                            double maxrange, int discretize, void *records, uint64_t capacity_records, uint64_t *counts);
int r3d_tree_apply_deltas_owned(r3d_tree *tree, const void *records, const uint64_t *counts, uint32_t n_scans, uint32_t part,
                                uint32_t nparts);
/* Same, but only NOTED: everything noted since the last flush -- every rank's share of a round of the multi-GPU merge -- is
 * applied by the next r3d_scan_deltas_compute call (or the first call that reads or updates the map, or
 * r3d_tree_flush_deferred) in ONE sorted, scan-ordered pass: owned records are keyed by (brick, position of the scan in
 * the round), radix-sorted, and every brick is updated by one warp in scan order -- three launches per round instead of one
 * per scan and rank.  `records` must stay valid until then; `counts` is copied.  Order among noted jobs and against later
 * applies is the call order.  R3D_ROUND_SORTED=0 applies them scan by scan instead. */
int r3d_tree_defer_deltas_owned(r3d_tree *tree, const void *records, const uint64_t *counts, uint32_t n_scans, uint32_t part, uint32_t nparts);
int r3d_tree_flush_deferred(r3d_tree *tree);
/* Expand delta records to explicit OcTreeKeys (n x 3 uint16) for inspection / parity tests (host buffers). */
int r3d_delta_expand_keys(const void *records_host, uint64_t n_records, uint16_t *free_keys, uint64_t free_cap,
                          uint64_t *n_free, uint16_t *occ_keys, uint64_t occ_cap, uint64_t *n_occ);

/* Statistics of the last scan delta: out[0] rays cast, out[1] free-cell visits (DDA steps), out[2] delta records,
 * out[3] bricks in the map as of the last counter read-back (r3d_tree_num_bricks is exact). */
int r3d_tree_last_scan_stats(r3d_tree *tree, uint64_t out[4]);
/* Host-side clock of the last pipelined r3d_tree_insert_scans batch, in nanoseconds: out[0] time the host spent blocked
 * on scan counters (the GPU was the slower side), out[1] time it spent queueing work, out[2] the longest single
 * turnaround between a scan's counters arriving and the next wait, out[3] scans that went through the pipeline. */
int r3d_tree_pipeline_stats(r3d_tree *tree, uint64_t out[4]);
/* Regrowth of the map while it was filled: out = { brick-pool regrowths (each copies the pool), hash-table regrowths (each
 * re-hashes), brick-pool capacity, hash-table capacity }.  r3d_tree_reserve() up front avoids them. */
int r3d_tree_growth_stats(r3d_tree *tree, uint64_t out[4]);

/* tree.updateInnerOccupancy() (octomap/txt_transfer_octomap.py:35): inner values are derived on demand. */
int r3d_tree_update_inner_occupancy(r3d_tree *tree);

/*
 * tree.writeBinary(path) (octomap/txt_transfer_octomap.py:36, octomap/ply_transfer_octomap.py:48):
 * toMaxLikelihood + prune + "# Octomap OcTree binary file" header + 2-bits-per-child pre-order stream.
 * The tree keeps its log-odds (upstream's writeBinary mutates them to the clamping values; use
 * r3d_tree_to_max_likelihood for that side effect).
 */
int r3d_tree_write_bt(r3d_tree *tree, const char *path);
/* Same into caller memory: header+payload; *len receives the size needed even when cap is too small. */
int r3d_tree_write_bt_mem(r3d_tree *tree, uint8_t *buf, size_t cap, size_t *len);
int r3d_tree_to_max_likelihood(r3d_tree *tree);
/*
 * tree.readBinary(path) of the `octomap` module (the step after the path: what octovis and the thesis' viewers do with
 * the .bt files of octomap/txt_transfer_octomap.py:36): replaces the tree's content (and resolution) with the file's;
 * occupied leaves get the upper clamping log-odds, free leaves the lower one, pruned leaves are expanded to voxels.
 */
int r3d_tree_read_bt(r3d_tree *tree, const char *path);
int r3d_tree_read_bt_mem(r3d_tree *tree, const uint8_t *data, size_t len);
/*
 * tree.write(path) / tree.read(path) of the `octomap` module: the .ot format, which keeps every node's log-odds
 * ("# Octomap OcTree file" header; pre-order, per node float32 value + one byte of child-exists bits).  The tree written
 * is upstream's tree as it stands after non-lazy updates (maximally pruned under exact equality of sibling leaf values,
 * inner values = max of the children).  Assembled on the host from the exported bricks (section 8f-4).
 */
int r3d_tree_write_ot(r3d_tree *tree, const char *path);
int r3d_tree_write_ot_mem(r3d_tree *tree, uint8_t *buf, size_t cap, size_t *len);
int r3d_tree_read_ot(r3d_tree *tree, const char *path);
int r3d_tree_read_ot_mem(r3d_tree *tree, const uint8_t *data, size_t len);

/* Queries used by tests / tools. */
int r3d_tree_num_voxels(r3d_tree *tree, uint64_t *n);            /* depth-16 leaves ever updated */
int r3d_tree_size(r3d_tree *tree, uint64_t *n_nodes);            /* upstream size(): nodes of the value-pruned tree */
int r3d_tree_search(r3d_tree *tree, const uint16_t *keys, uint64_t n, float *values, uint8_t *found);
int r3d_tree_export_voxels(r3d_tree *tree, uint16_t *keys, float *values, uint64_t cap, uint64_t *n);
/* coordToKeyChecked for n float32 points: keys n x 3 uint16, valid n bytes. */
int r3d_coord_to_key(r3d_tree *tree, const float *xyz, uint64_t n, uint16_t *keys, uint8_t *valid);

#ifdef __cplusplus
}
#endif
#endif /* R3D_H_ */
