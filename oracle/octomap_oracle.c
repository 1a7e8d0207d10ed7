/*
 * octomap_oracle.c -- CPU restatement of the OctoMap occupancy arithmetic used by the
 * reference scripts.  TEST INFRASTRUCTURE ONLY: nothing in the product path
 * (3d_reconstruction_system_b200/, transfer/, octomap/, other_tools/) may load this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED.  The arithmetic lives in a third-party dependency that is NOT in
 * /root/reference and is not version-pinned there (requirements.txt has no octomap
 * entry): the `octomap` Python extension (wkentaro/octomap-python or
 * neka-nat/python-octomap) wrapping OctoMap 1.8/1.9 C++.  Its published algorithm is
 * restated below from OcTreeBaseImpl.hxx (coordToKeyChecked, keyToCoord, search,
 * computeRayKeys, pruneNode, expandNode, prune), OccupancyOcTreeBase.hxx (updateNode,
 * updateNodeRecurs, updateNodeLogOdds, computeUpdate, insertPointCloud,
 * toMaxLikelihood, updateInnerOccupancy, writeBinaryNode) and
 * AbstractOccupancyOcTree.cpp (writeBinary / writeBinaryConst header text).
 * Reference call sites this anchors on:
 *   octomap/txt_transfer_octomap.py:25,33-36   OcTree(0.1), updateNode(p,True),
 *   octomap/ply_transfer_octomap.py:33,45-48   updateInnerOccupancy(), writeBinary()
 *   other_tools/ply_transfer_octomap.py:33,45-48 (identical copy)
 * The reference holds no tests / golden vectors for this boundary, so the known-answer
 * tests in tests/ are hand-derived from the .bt format.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off; no fast-math: float/double
 * operation order below is part of the contract).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TREE_DEPTH 16
#define TREE_MAX_VAL 32768

typedef struct Node {
    struct Node **children; /* NULL or array of 8 (entries may be NULL) */
    float value;            /* log-odds */
} Node;

typedef struct {
    Node *root;
    double resolution;
    double resolution_factor; /* 1.0 / resolution */
    size_t tree_size;
    float prob_hit_log, prob_miss_log, clamp_min, clamp_max, occ_thres_log;
} OTree;

/* ---- 48-bit key set (open addressing) used for the per-scan free / occupied sets ---- */
typedef struct {
    uint64_t *slots; /* key+1, 0 = empty */
    size_t cap, n;
} KeySet;

static uint64_t pack_key(const uint16_t k[3]) {
    return (uint64_t)k[0] | ((uint64_t)k[1] << 16) | ((uint64_t)k[2] << 32);
}
static void unpack_key(uint64_t p, uint16_t k[3]) {
    k[0] = (uint16_t)(p & 0xffff); k[1] = (uint16_t)((p >> 16) & 0xffff); k[2] = (uint16_t)((p >> 32) & 0xffff);
}
static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static void ks_init(KeySet *s, size_t cap_pow2) {
    s->cap = cap_pow2; s->n = 0; s->slots = (uint64_t *)calloc(cap_pow2, sizeof(uint64_t));
}
static void ks_free(KeySet *s) { free(s->slots); s->slots = NULL; s->cap = s->n = 0; }
static int ks_insert_raw(uint64_t *slots, size_t cap, uint64_t v) {
    size_t i = (size_t)mix64(v) & (cap - 1);
    for (;;) {
        if (slots[i] == 0) { slots[i] = v; return 1; }
        if (slots[i] == v) return 0;
        i = (i + 1) & (cap - 1);
    }
}
static void ks_grow(KeySet *s) {
    size_t ncap = s->cap * 2;
    uint64_t *ns = (uint64_t *)calloc(ncap, sizeof(uint64_t));
    for (size_t i = 0; i < s->cap; ++i) if (s->slots[i]) ks_insert_raw(ns, ncap, s->slots[i]);
    free(s->slots); s->slots = ns; s->cap = ncap;
}
static int ks_insert(KeySet *s, uint64_t packed) {
    if ((s->n + 1) * 2 > s->cap) ks_grow(s);
    int added = ks_insert_raw(s->slots, s->cap, packed + 1);
    s->n += (size_t)added;
    return added;
}
static int ks_contains(const KeySet *s, uint64_t packed) {
    uint64_t v = packed + 1;
    size_t i = (size_t)mix64(v) & (s->cap - 1);
    for (;;) {
        if (s->slots[i] == 0) return 0;
        if (s->slots[i] == v) return 1;
        i = (i + 1) & (s->cap - 1);
    }
}

/* ---- construction (OcTree(res): a9) ---- */
static float logodds(double p) { return (float)log(p / (1 - p)); }

OTree *oo_create(double resolution) {
    OTree *t = (OTree *)calloc(1, sizeof(OTree));
    t->root = NULL;
    t->resolution = resolution;
    t->resolution_factor = 1.0 / resolution;
    t->tree_size = 0;
    t->prob_hit_log = logodds(0.7);
    t->prob_miss_log = logodds(0.4);
    t->clamp_min = logodds(0.1192);
    t->clamp_max = logodds(0.971);
    t->occ_thres_log = logodds(0.5);
    return t;
}

static void free_node(Node *n) {
    if (!n) return;
    if (n->children) {
        for (int i = 0; i < 8; ++i) free_node(n->children[i]);
        free(n->children);
    }
    free(n);
}
void oo_clear(OTree *t) { free_node(t->root); t->root = NULL; t->tree_size = 0; }
void oo_destroy(OTree *t) { if (!t) return; oo_clear(t); free(t); }

void oo_params(const OTree *t, float out[5]) {
    out[0] = t->prob_hit_log; out[1] = t->prob_miss_log; out[2] = t->clamp_min; out[3] = t->clamp_max;
    out[4] = t->occ_thres_log;
}
size_t oo_size(const OTree *t) { return t->tree_size; }
double oo_resolution(const OTree *t) { return t->resolution; }

/* ---- keys (a10 step 2) ---- */
static int coord_to_key_checked1(const OTree *t, double coordinate, uint16_t *key) {
    int scaled = ((int)floor(t->resolution_factor * coordinate)) + TREE_MAX_VAL;
    if (scaled >= 0 && ((unsigned)scaled) < (2u * TREE_MAX_VAL)) { *key = (uint16_t)scaled; return 1; }
    return 0;
}
static int coord_to_key_checked3(const OTree *t, const float p[3], uint16_t key[3]) {
    for (int i = 0; i < 3; ++i) if (!coord_to_key_checked1(t, (double)p[i], &key[i])) return 0;
    return 1;
}
static double key_to_coord(const OTree *t, uint16_t key) {
    return ((double)((int)key - (int)TREE_MAX_VAL) + 0.5) * t->resolution;
}
int oo_coord_to_key(const OTree *t, double x, double y, double z, uint16_t key[3]) {
    float p[3] = {(float)x, (float)y, (float)z};
    return coord_to_key_checked3(t, p, key);
}
double oo_key_to_coord(const OTree *t, uint16_t key) { return key_to_coord(t, key); }

/* ---- node helpers ---- */
static Node *new_node(float v) { Node *n = (Node *)malloc(sizeof(Node)); n->children = NULL; n->value = v; return n; }
static int child_exists(const Node *n, unsigned i) { return n->children != NULL && n->children[i] != NULL; }
static int has_children(const Node *n) {
    if (!n->children) return 0;
    for (int i = 0; i < 8; ++i) if (n->children[i]) return 1;
    return 0;
}
static Node *create_child(OTree *t, Node *n, unsigned i) {
    if (!n->children) n->children = (Node **)calloc(8, sizeof(Node *));
    n->children[i] = new_node(0.0f);
    t->tree_size++;
    return n->children[i];
}
static void expand_node(OTree *t, Node *n) {
    for (unsigned i = 0; i < 8; ++i) { Node *c = create_child(t, n, i); c->value = n->value; }
}
static int is_collapsible(const Node *n) {
    if (!child_exists(n, 0)) return 0;
    const Node *first = n->children[0];
    if (has_children(first)) return 0;
    for (unsigned i = 1; i < 8; ++i) {
        if (!child_exists(n, i) || has_children(n->children[i]) || !(n->children[i]->value == first->value)) return 0;
    }
    return 1;
}
static int prune_node(OTree *t, Node *n) {
    if (!is_collapsible(n)) return 0;
    n->value = n->children[0]->value;
    for (unsigned i = 0; i < 8; ++i) { free_node(n->children[i]); t->tree_size--; }
    free(n->children); n->children = NULL;
    return 1;
}
static float max_child_logodds(const Node *n) {
    float m = -FLT_MAX;
    if (n->children) for (int i = 0; i < 8; ++i) if (n->children[i]) { float l = n->children[i]->value; if (l > m) m = l; }
    return m;
}
static unsigned child_idx(const uint16_t key[3], int depth_bit) {
    unsigned pos = 0;
    if (key[0] & (1 << depth_bit)) pos += 1;
    if (key[1] & (1 << depth_bit)) pos += 2;
    if (key[2] & (1 << depth_bit)) pos += 4;
    return pos;
}

static Node *search_key(const OTree *t, const uint16_t key[3]) {
    if (!t->root) return NULL;
    Node *cur = t->root;
    for (int i = TREE_DEPTH - 1; i >= 0; --i) {
        unsigned pos = child_idx(key, i);
        if (child_exists(cur, pos)) cur = cur->children[pos];
        else { if (!has_children(cur)) return cur; return NULL; }
    }
    return cur;
}

/* ---- updateNode (a10) ---- */
static void update_logodds(const OTree *t, Node *n, float update) {
    n->value += update;
    if (n->value < t->clamp_min) { n->value = t->clamp_min; return; }
    if (n->value > t->clamp_max) n->value = t->clamp_max;
}
static void update_recurs(OTree *t, Node *node, int just_created, const uint16_t key[3], unsigned depth, float update) {
    int created = 0;
    if (depth < TREE_DEPTH) {
        unsigned pos = child_idx(key, TREE_DEPTH - 1 - (int)depth);
        if (!child_exists(node, pos)) {
            if (!has_children(node) && !just_created) expand_node(t, node);
            else { create_child(t, node, pos); created = 1; }
        }
        update_recurs(t, node->children[pos], created, key, depth + 1, update);
        if (!prune_node(t, node)) node->value = max_child_logodds(node);
    } else {
        update_logodds(t, node, update);
    }
}
void oo_update_key_logodds(OTree *t, const uint16_t key[3], float update) {
    Node *leaf = search_key(t, key);
    if (leaf && ((update >= 0 && leaf->value >= t->clamp_max) || (update <= 0 && leaf->value <= t->clamp_min))) return;
    int created_root = 0;
    if (!t->root) { t->root = new_node(0.0f); t->tree_size++; created_root = 1; }
    update_recurs(t, t->root, created_root, key, 0, update);
}
void oo_update_key(OTree *t, const uint16_t key[3], int occupied) {
    oo_update_key_logodds(t, key, occupied ? t->prob_hit_log : t->prob_miss_log);
}
/* binding: updateNode(ndarray[double], bool) -> point3d(float,float,float) */
int oo_update_point(OTree *t, double x, double y, double z, int occupied) {
    uint16_t key[3];
    if (!oo_coord_to_key(t, x, y, z, key)) return 0;
    oo_update_key(t, key, occupied);
    return 1;
}
int oo_update_point_logodds(OTree *t, double x, double y, double z, float update) {
    uint16_t key[3];
    if (!oo_coord_to_key(t, x, y, z, key)) return 0;
    oo_update_key_logodds(t, key, update);
    return 1;
}
/* per-point loop of txt_read (octomap/txt_transfer_octomap.py:16-28); returns #points in range */
size_t oo_update_points(OTree *t, const double *xyz, size_t n, int occupied) {
    size_t ok = 0;
    for (size_t i = 0; i < n; ++i) ok += (size_t)oo_update_point(t, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], occupied);
    return ok;
}
size_t oo_update_points_f32(OTree *t, const float *xyz, size_t n, int occupied) {
    size_t ok = 0;
    for (size_t i = 0; i < n; ++i) {
        uint16_t key[3];
        if (!coord_to_key_checked3(t, &xyz[3 * i], key)) continue;
        oo_update_key(t, key, occupied);
        ok++;
    }
    return ok;
}

/* ---- computeRayKeys (a11) ---- */
/* returns number of keys written (<= max_keys), or -1 when origin/end is out of bounds */
static long ray_keys(const OTree *t, const float origin[3], const float end[3],
                     uint16_t *out, long max_keys, KeySet *set) {
    uint16_t key_origin[3], key_end[3];
    if (!coord_to_key_checked3(t, origin, key_origin) || !coord_to_key_checked3(t, end, key_end)) return -1;
    if (key_origin[0] == key_end[0] && key_origin[1] == key_end[1] && key_origin[2] == key_end[2]) return 0;
    long n = 0;
#define EMIT(K) do { if (set) ks_insert(set, pack_key(K)); \
        if (out && n < max_keys) { out[3*n] = (K)[0]; out[3*n+1] = (K)[1]; out[3*n+2] = (K)[2]; } n++; } while (0)
    EMIT(key_origin);

    float direction[3];
    for (int i = 0; i < 3; ++i) direction[i] = end[i] - origin[i];
    /* Vector3::norm(): sqrt of the float-evaluated sum of squares, returned as double */
    float nsq = direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2];
    float length = (float)sqrt((double)nsq);
    for (int i = 0; i < 3; ++i) direction[i] /= length;

    int step[3];
    double tMax[3], tDelta[3];
    uint16_t cur[3] = {key_origin[0], key_origin[1], key_origin[2]};
    for (int i = 0; i < 3; ++i) {
        if (direction[i] > 0.0) step[i] = 1;
        else if (direction[i] < 0.0) step[i] = -1;
        else step[i] = 0;
        if (step[i] != 0) {
            double voxelBorder = key_to_coord(t, cur[i]);
            voxelBorder += (float)(step[i] * t->resolution * 0.5);
            tMax[i] = (voxelBorder - origin[i]) / direction[i];
            tDelta[i] = t->resolution / fabs((double)direction[i]);
        } else {
            tMax[i] = DBL_MAX;
            tDelta[i] = DBL_MAX;
        }
    }
    for (;;) {
        unsigned dim;
        if (tMax[0] < tMax[1]) { if (tMax[0] < tMax[2]) dim = 0; else dim = 2; }
        else { if (tMax[1] < tMax[2]) dim = 1; else dim = 2; }
        cur[dim] = (uint16_t)(cur[dim] + step[dim]);
        tMax[dim] += tDelta[dim];
        if (cur[0] == key_end[0] && cur[1] == key_end[1] && cur[2] == key_end[2]) break;
        double d01 = tMax[0] < tMax[1] ? tMax[0] : tMax[1];
        double dist_from_origin = d01 < tMax[2] ? d01 : tMax[2];
        if (dist_from_origin > length) break;
        EMIT(cur);
    }
#undef EMIT
    return n;
}
long oo_compute_ray_keys(const OTree *t, const float origin[3], const float end[3], uint16_t *out, long max_keys) {
    return ray_keys(t, origin, end, out, max_keys, NULL);
}

/* ---- computeUpdate / insertPointCloud (a11) ---- */
static double vnorm3(const float v[3]) {
    float nsq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    return sqrt((double)nsq);
}
static void compute_update(const OTree *t, const float *pts, size_t n, const float origin[3], double maxrange,
                           KeySet *free_cells, KeySet *occ_cells) {
    for (size_t i = 0; i < n; ++i) {
        const float *p = &pts[3 * i];
        float d[3] = {p[0] - origin[0], p[1] - origin[1], p[2] - origin[2]};
        if (maxrange < 0.0 || vnorm3(d) <= maxrange) {
            ray_keys(t, origin, p, NULL, 0, free_cells);
            uint16_t key[3];
            if (coord_to_key_checked3(t, p, key)) ks_insert(occ_cells, pack_key(key));
        } else {
            /* direction = (p - origin).normalized(); new_end = origin + direction * (float)maxrange */
            double len = vnorm3(d);
            if (len > 0) { float fl = (float)len; d[0] /= fl; d[1] /= fl; d[2] /= fl; }
            float mr = (float)maxrange;
            float new_end[3] = {origin[0] + d[0] * mr, origin[1] + d[1] * mr, origin[2] + d[2] * mr};
            ray_keys(t, origin, new_end, NULL, 0, free_cells);
        }
    }
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static size_t ks_dump_sorted(const KeySet *s, const KeySet *exclude, uint64_t **out) {
    uint64_t *v = (uint64_t *)malloc((s->n ? s->n : 1) * sizeof(uint64_t));
    size_t m = 0;
    for (size_t i = 0; i < s->cap; ++i) if (s->slots[i]) {
        uint64_t k = s->slots[i] - 1;
        if (exclude && ks_contains(exclude, k)) continue;
        v[m++] = k;
    }
    qsort(v, m, sizeof(uint64_t), cmp_u64);
    *out = v;
    return m;
}
/* discretize=True pre-pass (computeDiscreteUpdate): one voxel-centre point per distinct endpoint key */
static size_t discretize_points(const OTree *t, const float *pts, size_t n, float **out) {
    KeySet seen; ks_init(&seen, 1024);
    float *d = (float *)malloc((n ? n : 1) * 3 * sizeof(float));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        uint16_t key[3];
        /* upstream uses coordToKey (unchecked cast); out-of-range points are kept as-is by wrapping -- we
           restate the checked behaviour and drop them, which is what the checked ray/endpoint code does next */
        if (!coord_to_key_checked3(t, &pts[3 * i], key)) continue;
        if (ks_insert(&seen, pack_key(key))) {
            for (int a = 0; a < 3; ++a) d[3 * m + a] = (float)key_to_coord(t, key[a]);
            m++;
        }
    }
    ks_free(&seen);
    *out = d;
    return m;
}

/* Computes the per-scan free / occupied key sets (free already minus occupied), sorted by packed key.
 * Caller frees with oo_free. */
void oo_compute_update(const OTree *t, const float *pts, size_t n, const float origin[3], double maxrange,
                       uint64_t **free_out, size_t *n_free, uint64_t **occ_out, size_t *n_occ) {
    KeySet fr, oc; ks_init(&fr, 1 << 16); ks_init(&oc, 1 << 12);
    compute_update(t, pts, n, origin, maxrange, &fr, &oc);
    *n_free = ks_dump_sorted(&fr, &oc, free_out);
    *n_occ = ks_dump_sorted(&oc, NULL, occ_out);
    ks_free(&fr); ks_free(&oc);
}
void oo_free(void *p) { free(p); }

void oo_insert_point_cloud_f32(OTree *t, const float *pts, size_t n, const float origin[3], double maxrange,
                               int discretize) {
    float *dpts = NULL;
    if (discretize) { n = discretize_points(t, pts, n, &dpts); pts = dpts; }
    KeySet fr, oc; ks_init(&fr, 1 << 16); ks_init(&oc, 1 << 12);
    compute_update(t, pts, n, origin, maxrange, &fr, &oc);
    uint16_t key[3];
    for (size_t i = 0; i < fr.cap; ++i) if (fr.slots[i]) {
        uint64_t k = fr.slots[i] - 1;
        if (ks_contains(&oc, k)) continue; /* occupied wins */
        unpack_key(k, key); oo_update_key(t, key, 0);
    }
    for (size_t i = 0; i < oc.cap; ++i) if (oc.slots[i]) { unpack_key(oc.slots[i] - 1, key); oo_update_key(t, key, 1); }
    ks_free(&fr); ks_free(&oc);
    free(dpts);
}
/* binding: insertPointCloud(ndarray[double,N,3], ndarray[double,3] origin, maxrange, lazy_eval, discretize) */
void oo_insert_point_cloud(OTree *t, const double *xyz, size_t n, const double origin[3], double maxrange,
                           int discretize) {
    float *p = (float *)malloc((n ? n : 1) * 3 * sizeof(float));
    for (size_t i = 0; i < 3 * n; ++i) p[i] = (float)xyz[i];
    float o[3] = {(float)origin[0], (float)origin[1], (float)origin[2]};
    oo_insert_point_cloud_f32(t, p, n, o, maxrange, discretize);
    free(p);
}

/* ---- updateInnerOccupancy (a12) ---- */
static void update_inner_recurs(Node *n, unsigned depth) {
    if (has_children(n)) {
        if (depth < TREE_DEPTH) for (int i = 0; i < 8; ++i) if (child_exists(n, (unsigned)i)) update_inner_recurs(n->children[i], depth + 1);
        n->value = max_child_logodds(n);
    }
}
void oo_update_inner_occupancy(OTree *t) { if (t->root) update_inner_recurs(t->root, 0); }

/* ---- queries ---- */
int oo_search(const OTree *t, const uint16_t key[3], float *value) {
    Node *n = search_key(t, key);
    if (!n) return 0;
    if (value) *value = n->value;
    return 1;
}
static void leaves_recurs(const Node *n, unsigned depth, uint16_t kx, uint16_t ky, uint16_t kz,
                          uint16_t *keys, float *vals, uint8_t *depths, size_t cap, size_t *cnt) {
    if (!has_children(n)) {
        if (*cnt < cap) {
            if (keys) { keys[3 * *cnt] = kx; keys[3 * *cnt + 1] = ky; keys[3 * *cnt + 2] = kz; }
            if (vals) vals[*cnt] = n->value;
            if (depths) depths[*cnt] = (uint8_t)depth;
        }
        (*cnt)++;
        return;
    }
    int bit = TREE_DEPTH - 1 - (int)depth;
    for (unsigned i = 0; i < 8; ++i) if (child_exists(n, i))
        leaves_recurs(n->children[i], depth + 1, (uint16_t)(kx | ((i & 1) << bit)), (uint16_t)(ky | (((i >> 1) & 1) << bit)),
                      (uint16_t)(kz | (((i >> 2) & 1) << bit)), keys, vals, depths, cap, cnt);
}
/* leaf iteration in child-index (pre-order) order; keys are the minimum-corner key of the leaf's cube */
size_t oo_leaves(const OTree *t, uint16_t *keys, float *vals, uint8_t *depths, size_t cap) {
    size_t cnt = 0;
    if (t->root) leaves_recurs(t->root, 0, 0, 0, 0, keys, vals, depths, cap, &cnt);
    return cnt;
}

/* ---- writeBinary (a13) ---- */
static void ml_node(const OTree *t, Node *n) { n->value = (n->value >= t->occ_thres_log) ? t->clamp_max : t->clamp_min; }
static void ml_recurs(const OTree *t, Node *n, unsigned depth, unsigned max_depth) {
    if (depth < max_depth) { for (unsigned i = 0; i < 8; ++i) if (child_exists(n, i)) ml_recurs(t, n->children[i], depth + 1, max_depth); }
    else ml_node(t, n);
}
void oo_to_max_likelihood(OTree *t) {
    if (!t->root) return;
    for (unsigned depth = TREE_DEPTH; depth > 0; depth--) ml_recurs(t, t->root, 0, depth);
    ml_node(t, t->root);
}
static void prune_recurs(OTree *t, Node *n, unsigned depth, unsigned max_depth, unsigned *num_pruned) {
    if (depth < max_depth) { for (unsigned i = 0; i < 8; ++i) if (child_exists(n, i)) prune_recurs(t, n->children[i], depth + 1, max_depth, num_pruned); }
    else if (prune_node(t, n)) (*num_pruned)++;
}
void oo_prune(OTree *t) {
    if (!t->root) return;
    for (unsigned depth = TREE_DEPTH - 1; depth > 0; --depth) {
        unsigned num_pruned = 0;
        prune_recurs(t, t->root, 0, depth, &num_pruned);
        if (num_pruned == 0) break;
    }
}
typedef struct { uint8_t *buf; size_t len, cap; } ByteBuf;
static void bb_put(ByteBuf *b, const void *p, size_t n) {
    if (b->len + n > b->cap) { while (b->len + n > b->cap) b->cap = b->cap ? b->cap * 2 : 4096; b->buf = (uint8_t *)realloc(b->buf, b->cap); }
    memcpy(b->buf + b->len, p, n); b->len += n;
}
static void write_binary_node(const OTree *t, const Node *n, ByteBuf *b) {
    uint8_t c[2] = {0, 0};
    for (unsigned i = 0; i < 8; ++i) {
        if (!child_exists(n, i)) continue;
        const Node *ch = n->children[i];
        unsigned sh = (i & 3) * 2;
        if (has_children(ch)) c[i >> 2] |= (uint8_t)(3u << sh);
        else if (ch->value >= t->occ_thres_log) c[i >> 2] |= (uint8_t)(2u << sh); /* bit 2i+1: occupied */
        else c[i >> 2] |= (uint8_t)(1u << sh);                                     /* bit 2i  : free */
    }
    bb_put(b, c, 2);
    for (unsigned i = 0; i < 8; ++i) if (child_exists(n, i) && has_children(n->children[i])) write_binary_node(t, n->children[i], b);
}
/* ostream << double with default precision 6 == "%g" */
static void fmt_res(double r, char *out, size_t n) { snprintf(out, n, "%g", r); }

/* writeBinary(): toMaxLikelihood(); prune(); header; data.  Mutates the tree like upstream.
 * Returns a malloc'd buffer (free with oo_free) and its length. */
uint8_t *oo_write_binary_mem(OTree *t, size_t *len) {
    oo_to_max_likelihood(t);
    oo_prune(t);
    ByteBuf b = {NULL, 0, 0};
    char hdr[512], res[64];
    fmt_res(t->resolution, res, sizeof res);
    int hl = snprintf(hdr, sizeof hdr,
                      "# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n"
                      "id OcTree\nsize %zu\nres %s\ndata\n", t->tree_size, res);
    bb_put(&b, hdr, (size_t)hl);
    if (t->root) write_binary_node(t, t->root, &b);
    *len = b.len;
    return b.buf;
}
/* write() (AbstractOcTree::write + OcTreeBaseImpl::writeData / writeNodesRecurs): the full tree, log-odds preserved.
 * Header like .bt but for the first line "# Octomap OcTree file"; then pre-order, per node: float32 value, one byte with
 * bit i set when child i exists, then the existing children in index order.  The tree is written as it is (no
 * max-likelihood conversion, no extra pruning). */
static void write_ot_node(const Node *n, ByteBuf *b) {
    bb_put(b, &n->value, sizeof(float));
    unsigned char mask = 0;
    for (unsigned i = 0; i < 8; ++i) if (child_exists(n, i)) mask |= (unsigned char)(1u << i);
    bb_put(b, &mask, 1);
    for (unsigned i = 0; i < 8; ++i) if (child_exists(n, i)) write_ot_node(n->children[i], b);
}
uint8_t *oo_write_ot_mem(const OTree *t, size_t *len) {
    ByteBuf b = {NULL, 0, 0};
    char hdr[512], res[64];
    fmt_res(t->resolution, res, sizeof res);
    int hl = snprintf(hdr, sizeof hdr,
                      "# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n"
                      "id OcTree\nsize %zu\nres %s\ndata\n", t->tree_size, res);
    bb_put(&b, hdr, (size_t)hl);
    if (t->root) write_ot_node(t->root, &b);
    *len = b.len;
    return b.buf;
}

int oo_write_binary(OTree *t, const char *path) {
    size_t len; uint8_t *buf = oo_write_binary_mem(t, &len);
    FILE *f = fopen(path, "wb");
    if (!f) { free(buf); return 0; }
    size_t w = fwrite(buf, 1, len, f);
    fclose(f); free(buf);
    return w == len;
}
