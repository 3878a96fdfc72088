/*
 * cantucci_oracle.h -- CPU oracle for the cantucci hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is a plain-C restatement of the reference's distance estimator and
 * naive-surface-nets mesher, written from the Rust sources under
 * /root/reference (file:line cited at every function in cantucci_oracle.c).
 *
 * PARITY UNPINNED: the reference ships no golden vectors, no #[test] and no
 * expected outputs for this path, and it cannot be
 * compiled here (no rustc/cargo).  The oracle is therefore pinned only by
 * (i) an independent numpy-float32 restatement (tests/test_oracle_numpy.py),
 * (ii) analytic checks on the Sphere shape, (iii) frozen golden vectors made
 * by THIS oracle (tests/golden/, generator committed).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (cantucci_b200/) never links, imports or calls it.
 */
#ifndef CANTUCCI_ORACLE_H
#define CANTUCCI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Range<Point3<f32>>  (src/octree/mod.rs:13) */
typedef struct { float start[3]; float end[3]; } orc_span;

/* mesh::Vertex  (src/mesh/mod.rs:255-261), #[repr(C)], 28 bytes */
typedef struct { float position[3]; float normal[3]; float distance_from_surface; } orc_vertex;

enum { ORC_SHAPE_MANDELBULB = 0, ORC_SHAPE_SPHERE = 1 };

/* Mandelbulb<P>{max_iters,bailout} (src/shape/mandelbulb.rs:13-16) or
 * Sphere{center,radius} (src/shape/sphere.rs:7-10). */
typedef struct {
    int32_t  kind;
    uint32_t power;        /* const generic P */
    uint64_t max_iters;
    float    bailout;
    float    center[3];    /* sphere */
    float    radius;       /* sphere */
} orc_shape;

/* Extra per-sample facts used for flop accounting and for the tolerance
 * filter ("away from the escape boundary"). */
typedef struct {
    uint32_t iters;        /* completed iterations k (rotate executed) */
    uint32_t bailed;       /* 1 if the loop left through `r > bailout` */
    float    r;            /* r at exit (magnitude before the last rotate) */
    float    dr;
    float    min_margin;   /* min over iterations of |r - bailout| / bailout */
} orc_de_info;

typedef struct {
    orc_vertex *vertices;
    uint32_t   *indices;
    uint64_t    n_vertices;
    uint64_t    n_indices;
    double      first_s, second_s, third_s;  /* Timings (src/mesh/buffer.rs:398-405) */
    int32_t     panicked;   /* 1 = the reference would have panicked (lerp assert, math.rs:19) */
} orc_mesh;

float orc_min_distance_from(const orc_shape *s, const float p[3]);
float orc_min_distance_from_info(const orc_shape *s, const float p[3], orc_de_info *info);
void  orc_batch_min_distance_from(const orc_shape *s, const float *xyz, size_t n, float *out);

/* rotate variants exposed for the trig-vs-polynomial check */
void orc_rotate_p8_scalar(const float in[3], float out[3]);
void orc_rotate_generic(uint32_t power, const float in[3], float out[3]);
void orc_rotate(uint32_t power, const float in[3], float out[3]);

/* glibc 2.39 logf restated (FMA-contracted variant == __logf_fma). */
float orc_logf_glibc_fma(float x);

/* pass 1 only: (R+1)^3 samples of the EXPANDED span, x-major z-fastest */
int  orc_sample_grid(const orc_shape *s, const orc_span *span, uint32_t resolution, float *out);
/* optional: iteration histogram accumulation for flop accounting */
int  orc_sample_grid_info(const orc_shape *s, const orc_span *span, uint32_t resolution,
                          float *out, uint64_t *iter_hist /* max_iters+1 */, uint64_t *n_bailed);

int  orc_sample_grid_iters(const orc_shape *s, const orc_span *span, uint32_t resolution,
                           float *out, uint8_t *iters_out);

/* MeshBuffer::generate_for_box (src/mesh/buffer.rs:30-42). Returns 0 ok,
 * 1 = assertion on arguments would fire. out must be released with orc_mesh_free. */
int  orc_generate_for_box(const orc_shape *s, const orc_span *span, uint32_t resolution, orc_mesh *out);
void orc_mesh_free(orc_mesh *m);

/* Thread-pool driver mirroring mesh/mod.rs:61-62,141-148: one job per span on
 * nthreads workers.  meshes[nspans] is filled; returns wall seconds. */
double orc_generate_for_boxes_mt(const orc_shape *s, const orc_span *spans, size_t nspans,
                                 uint32_t resolution, int nthreads, orc_mesh *meshes);
/* Same, and also the sign bit-plane of every span's sample grid: planes[nspans][((R+1)^3 + 31) / 32],
 * bit j = !is_sign_positive(dists[j]).  Used by the parity gate to count sign mismatches. */
double orc_generate_for_boxes_signs_mt(const orc_shape *s, const orc_span *spans, size_t nspans,
                                       uint32_t resolution, int nthreads, orc_mesh *meshes, uint32_t *planes);
/* Same pool, pass 1 only (samples), results discarded except a checksum. */
double orc_sample_grids_mt(const orc_shape *s, const orc_span *spans, size_t nspans,
                           uint32_t resolution, int nthreads, double *checksum);

/* octree span maths (src/octree/mod.rs:21-23, 315-329) */
void orc_span_center(const orc_span *s, float out[3]);
void orc_create_spans(const orc_span *parent, orc_span out[8]);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
