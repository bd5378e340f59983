/*
 * oracle/restate.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, tape-free CPU restatement of the reference's hot path (forward path
 * tracer + reverse-mode gradient), one function per reference function, each
 * citing the file:line in /root/reference it follows.  It exists to check the
 * CUDA path on machines where /root/reference is absent (the GPU box) and to
 * serve as bench.py's `cpu_baseline` of kind "port".
 *
 * Pinning: the reference ships no tests, so this file is pinned
 * against (a) the UNMODIFIED reference headers compiled by oracle/Makefile into
 * oracle/_ref/libdrt_ref.so (tests/test_oracle.py::test_restatement_vs_reference,
 * run wherever /root/reference exists), (b) the golden vectors that library
 * produced, committed under tests/golden/ with their generator
 * (tests/golden/make_golden.py), and (c) the survey's known-answer vectors
 * KAT-1..4.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libdrtb.so) never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "drtb.h"

/* ---- sample stream (drtb.h; feeds include/drt/random.hpp:7-10) ----------- */

static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

typedef struct { uint64_t key; uint32_t ctr; } stream_t;

/* random::uniform(), random.hpp:7-10: double(rand()) / RAND_MAX */
static double uniform(stream_t* s)
{
    uint64_t k = splitmix64(s->key * 0x100000001B3ull + s->ctr++) % 2147483647ull;
    return (double)k / 2147483647.0;
}

/* ---- Vector<T,3> helpers, vector.hpp:573-600 ------------------------------ */

typedef struct { double x, y, z; } v3;

/* dot(): elementwise product then accumulate from T(), vector.hpp:573-578 */
static double dot3(v3 a, v3 b) { return ((0.0 + a.x*b.x) + a.y*b.y) + a.z*b.z; }
static v3 add3(v3 a, v3 b) { v3 r = {a.x+b.x, a.y+b.y, a.z+b.z}; return r; }
static v3 sub3(v3 a, v3 b) { v3 r = {a.x-b.x, a.y-b.y, a.z-b.z}; return r; }
static v3 mul3(v3 a, double s) { v3 r = {a.x*s, a.y*s, a.z*s}; return r; }
static v3 div3(v3 a, double s) { v3 r = {a.x/s, a.y/s, a.z/s}; return r; }
/* normalize(): v / sqrt(dot(v,v)), vector.hpp:580-590 */
static v3 normalize3(v3 a) { return div3(a, sqrt(dot3(a, a))); }
/* cross(), vector.hpp:592-600 */
static v3 cross3(v3 a, v3 b)
{
    v3 r = {a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x};
    return r;
}

static const double PI = 3.14159265358979323846;      /* constants.hpp:9 */

/* ---- shapes, shape.hpp ----------------------------------------------------- */

/* Plane::intersect, shape.hpp:49-56 */
static int plane_intersect(const drtb_prim* p, v3 o, v3 d, double* t)
{
    v3 n = {p->v[0], p->v[1], p->v[2]};
    v3 nn = {-1.0*n.x, -1.0*n.y, -1.0*n.z};            /* operator-: -1*v, vector.hpp:320-325 */
    double h = dot3(o, n) - p->v[3];
    *t = h / dot3(d, nn);
    return *t > 0;
}

/* Sphere::intersect, shape.hpp:78-103 (a is hard-coded to 1, :83) */
static int sphere_intersect(const drtb_prim* p, v3 o, v3 d, double* t)
{
    v3 c = {p->v[0], p->v[1], p->v[2]};
    double r = p->v[3];
    v3 oc = sub3(o, c);
    double a = 1;
    double b = 2 * dot3(oc, d);
    double cc = dot3(oc, oc) - r*r;
    double disc = b*b - 4*a*cc;
    if (disc < 0) return 0;
    double t1 = (-b - sqrt(disc)) / (2 * a);
    double t2 = (-b + sqrt(disc)) / (2 * a);
    if (t1 > 0 && t2 > 0) { *t = t1 < t2 ? t1 : t2; return 1; }
    if (t1 > 0) { *t = t1; return 1; }
    if (t2 > 0) { *t = t2; return 1; }
    return 0;
}

/* Triangle (NEW, no reference counterpart; semantics fixed in drtb.h):
 * Moller-Trumbore, both sides, inclusive barycentric bounds, t > 0. */
static v3 mesh_vertex(const drtb_mesh* m, int64_t tri, int corner)
{
    const double* p = m->vertices + 3 * (int64_t)m->indices[3 * tri + corner];
    v3 r = {p[0], p[1], p[2]};
    return r;
}

static int triangle_intersect(const drtb_mesh* m, int64_t tri, v3 o, v3 d, double* t)
{
    v3 v0 = mesh_vertex(m, tri, 0);
    v3 e1 = sub3(mesh_vertex(m, tri, 1), v0), e2 = sub3(mesh_vertex(m, tri, 2), v0);
    v3 p = cross3(d, e2);
    double det = dot3(e1, p);
    if (det == 0.0) return 0;
    double inv = 1.0 / det;
    v3 tv = sub3(o, v0);
    double u = dot3(tv, p) * inv;
    if (!(u >= 0.0 && u <= 1.0)) return 0;
    v3 q = cross3(tv, e1);
    double v = dot3(d, q) * inv;
    if (!(v >= 0.0 && u + v <= 1.0)) return 0;
    *t = dot3(e2, q) * inv;
    return *t > 0;
}

static v3 triangle_normal(const drtb_mesh* m, int64_t tri)
{
    v3 v0 = mesh_vertex(m, tri, 0);
    return normalize3(cross3(sub3(mesh_vertex(m, tri, 1), v0), sub3(mesh_vertex(m, tri, 2), v0)));
}

/* Pathtracer::raycast, pathtracer.hpp:72-89: linear scan in scene order,
 * `!hit || t >= tmin -> continue`, so the first shape wins ties.  Triangles
 * follow the analytic primitives in scene order (index n_prims + i). */
static int64_t raycast(const drtb_scene* s, const drtb_mesh* mesh, v3 o, v3 d, v3* point, v3* normal)
{
    double tmin = INFINITY;
    int64_t best = -1;
    for (int i = 0; i < s->n_prims; ++i) {
        const drtb_prim* p = &s->prims[i];
        double t;
        int hit = p->type == DRTB_PLANE ? plane_intersect(p, o, d, &t)
                                        : sphere_intersect(p, o, d, &t);
        if (!hit || t >= tmin) continue;
        tmin = t;
        best = i;
        *point = add3(o, mul3(d, t));
        if (p->type == DRTB_PLANE) {                     /* shape.hpp:58-59, RAW */
            v3 n = {p->v[0], p->v[1], p->v[2]};
            *normal = n;
        } else {                                         /* shape.hpp:105-106 */
            v3 c = {p->v[0], p->v[1], p->v[2]};
            *normal = normalize3(sub3(*point, c));
        }
    }
    if (mesh)
        for (int64_t i = 0; i < mesh->n_triangles; ++i) {
            double t;
            if (!triangle_intersect(mesh, i, o, d, &t) || t >= tmin) continue;
            tmin = t;
            best = s->n_prims + i;
            *point = add3(o, mul3(d, t));
            *normal = triangle_normal(mesh, i);
        }
    return isinf(tmin) ? -1 : best;
}

/* material lookup by scene index: analytic prims through materials[], triangles direct */
static int prim_color(const drtb_scene* s, const drtb_mesh* m, int64_t k)
{
    if (k < s->n_prims) return s->prims[k].material >= 0 ? s->materials[s->prims[k].material].color : -1;
    return m->color ? m->color[k - s->n_prims] : -1;
}
/* SpecularBxDF exponent of an analytic primitive, or -1 for DiffuseBxDF (triangles are diffuse, drtb.h) */
static double prim_specular(const drtb_scene* s, int64_t k)
{
    if (k >= s->n_prims || s->prims[k].material < 0) return -1.0;
    const drtb_material* m = &s->materials[s->prims[k].material];
    return m->type == DRTB_SPECULAR ? m->exponent : -1.0;
}
static int prim_emission(const drtb_scene* s, const drtb_mesh* m, int64_t k)
{
    if (k < s->n_prims) return s->prims[k].emission;
    return m->emission ? m->emission[k - s->n_prims] : -1;
}

/* make_frame, bxdf.hpp:29-41 (normal used RAW) */
static void make_frame(v3 n, v3* tangent, v3* bitangent)
{
    v3 e1 = {1, 0, 0}, e2 = {0, 1, 0};
    if (fabs(dot3(e1, n)) < fabs(dot3(e2, n)))
        *tangent = normalize3(sub3(e1, mul3(n, dot3(e1, n))));
    else
        *tangent = normalize3(sub3(e2, mul3(n, dot3(e2, n))));
    *bitangent = normalize3(cross3(n, *tangent));
}

/* angle_to_dir, bxdf.hpp:43-52 */
static v3 angle_to_dir(double theta, double phi, v3 tangent, v3 bitangent, v3 n)
{
    double x = cos(phi) * sin(theta);
    double y = sin(phi) * sin(theta);
    double z = cos(theta);
    return add3(add3(mul3(tangent, x), mul3(bitangent, y)), mul3(n, z));
}

/* ---- DiffuseBxDF::sample, bxdf.hpp:69-79 ---------------------------------- */
static v3 diffuse_sample(stream_t* rng, v3 n, double* pdf)
{
    double theta = asin(sqrt(uniform(rng)));
    double phi = 2 * PI * uniform(rng);
    v3 tangent, bitangent;
    make_frame(n, &tangent, &bitangent);
    v3 dir = angle_to_dir(theta, phi, tangent, bitangent, n);
    *pdf = cos(theta) / PI;
    return dir;
}

/* reflect(), vector.hpp:602-606: -v + 2*dot(n, v)*n */
static v3 reflect3(v3 v, v3 n)
{
    return add3(mul3(v, -1.0), mul3(n, 2 * dot3(n, v)));
}

/* ---- SpecularBxDF::sample, bxdf.hpp:107-120 -------------------------------- */
static v3 specular_sample(stream_t* rng, v3 n, v3 dir_in, double e, double* pdf)
{
    double theta = acos(sqrt(pow(uniform(rng), 2 / (e + 2))));
    double phi = 2 * PI * uniform(rng);
    v3 tangent, bitangent;
    make_frame(n, &tangent, &bitangent);
    v3 halfway = angle_to_dir(theta, phi, tangent, bitangent, n);
    if (dot3(halfway, dir_in) < 0)
        halfway = reflect3(halfway, n);
    v3 dir = reflect3(dir_in, halfway);
    *pdf = (e + 2) / (2 * PI) * pow(cos(theta), e + 1) * sin(theta);
    return dir;
}

/* ---- SpecularBxDF::operator(), bxdf.hpp:93-105: the scalar in front of m_color */
static double specular_factor(v3 n, v3 dir_in, v3 dir_out, double e)
{
    v3 halfway = normalize3(add3(dir_in, dir_out));
    double cos_theta = dot3(n, halfway);
    double sin_theta = sqrt(1 - cos_theta * cos_theta);
    return (e + 2) / (2 * PI) * pow(cos_theta, e) * sin_theta;
}

/* ---- one path: Pathtracer::trace/scatter, pathtracer.hpp:91-136, unrolled
 *      from recursion into a vertex list, then the tape's backward
 *      (vector.hpp:418-486) as two sweeps -------------- */

/* is_spec == 0: DiffuseBxDF (brdf = color/pi); else SpecularBxDF with brdf = spec * color */
typedef struct { int64_t prim; double p, cosn, pdf, spec; int is_spec; } vertex_t;

typedef struct {
    vertex_t* v; int n, cap;
    double* L;                       /* (n+1) x 3 suffix radiances */
    int Lcap;
} path_buf;

static void push_vertex(path_buf* b, vertex_t vx)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 64;
        b->v = (vertex_t*)realloc(b->v, sizeof(vertex_t) * (size_t)b->cap);
    }
    b->v[b->n++] = vx;
}

typedef struct { uint64_t segments, lit; } counters_t;

/* Traces from (o, d) with the stream positioned after the camera draws;
 * returns L_0 in out[3]; if g0 != NULL accumulates g0 . dL/dparam into grad. */
static void trace_path(const drtb_scene* s, const drtb_mesh* mesh, const drtb_render_opts* opt,
                       stream_t* rng, v3 o, v3 d, path_buf* b,
                       double out[3], const double* g0, double* grad,
                       counters_t* cnt)
{
    const int mb = opt->min_bounces;
    const double absorb = opt->absorb;
    b->n = 0;
    for (int depth = 0;; ++depth) {
        double p = 1;
        if (depth >= mb) {                               /* pathtracer.hpp:128-130 */
            if (uniform(rng) < absorb) break;
            p = 1 - absorb;
        }
        v3 pt = {0, 0, 0}, n = {0, 0, 0};
        int64_t k = raycast(s, mesh, o, d, &pt, &n);
        cnt->segments++;
        if (k < 0) break;                                /* :134-135 */
        vertex_t vx = {k, p, 0.0, 1.0, 0.0, 0};
        if (prim_color(s, mesh, k) < 0) {
            /* null BxDF: dir_out = 0, pdf = 1, brdf = 0 (pathtracer.hpp:25-26,
             * 38-39): every deeper term is multiplied by 0, so the path ends
             * here for all observable purposes. */
            push_vertex(b, vx);
            break;
        }
        double pdf;
        v3 dout;
        const double e = prim_specular(s, k);
        if (e >= 0) {
            v3 dir_in = mul3(d, -1.0);                   /* -dir_in, :101, 109 */
            dout = specular_sample(rng, n, dir_in, e, &pdf);
            vx.spec = specular_factor(n, dir_in, dout, e);   /* eval_bxdf, :100-101 */
            vx.is_spec = 1;
        } else {
            dout = diffuse_sample(rng, n, &pdf);         /* :106-109 */
        }
        vx.cosn = dot3(n, dout);                         /* :103 */
        vx.pdf = pdf;
        push_vertex(b, vx);
        o = add3(pt, mul3(dout, 1e-3));                  /* :99 */
        d = dout;
    }

    /* backward sweep: L_v = (E_v + (0 + ((rho_v/pi * L_{v+1}) * cos) / pdf)) / p_v
     * in exactly the op order of pathtracer.hpp:100-104,114,133 and
     * integrate.hpp:31-36. */
    const int n = b->n;
    if ((n + 1) * 3 > b->Lcap) {
        b->Lcap = (n + 1) * 3 + 192;
        b->L = (double*)realloc(b->L, sizeof(double) * (size_t)b->Lcap);
    }
    double* L = b->L;
    L[3*n] = L[3*n+1] = L[3*n+2] = 0.0;
    for (int v = n - 1; v >= 0; --v) {
        const vertex_t* vx = &b->v[v];
        const int em = prim_emission(s, mesh, vx->prim), colp = prim_color(s, mesh, vx->prim);
        for (int c = 0; c < 3; ++c) {
            double E = em >= 0 ? s->params[3*em + c] : 0.0;
            double diffuse = 0.0;
            if (colp >= 0) {
                double rho = s->params[3*colp + c];
                double brdf = vx->is_spec ? vx->spec * rho : rho / PI;
                diffuse = 0.0 + ((brdf * L[3*(v+1)+c]) * vx->cosn) / vx->pdf;
            }
            L[3*v+c] = (E + diffuse) / vx->p;
        }
    }
    out[0] = L[0]; out[1] = L[1]; out[2] = L[2];
    if (L[0] != 0.0 || L[1] != 0.0 || L[2] != 0.0) cnt->lit++;

    /* forward adjoint sweep == the tape's recursive backward, vector.hpp:418-486 */
    if (!g0 || !grad) return;
    double g[3] = {g0[0], g0[1], g0[2]};
    for (int v = 0; v < n; ++v) {
        const vertex_t* vx = &b->v[v];
        const int em = prim_emission(s, mesh, vx->prim), colp = prim_color(s, mesh, vx->prim);
        for (int c = 0; c < 3; ++c) {
            double gp = g[c] / vx->p;                    /* ScalarDivBackward (/p) */
            if (em >= 0)                                 /* AddBackward -> VariableNode */
                grad[3*em + c] += gp;
            if (colp >= 0) {
                int col = colp;
                double g2 = gp / vx->pdf;                /* ScalarDivBackward (/pdf) */
                double g3 = vx->cosn * g2;               /* ScalarMulBackward (*cos) */
                const int is_spec = vx->is_spec;
                /* MulBackward lhs: brdf.backward(radiance * g3); brdf = color/pi or factor*color */
                grad[3*col + c] += is_spec ? vx->spec * (L[3*(v+1)+c] * g3) : (L[3*(v+1)+c] * g3) / PI;
                /* MulBackward rhs: radiance.backward(brdf * g3) */
                g[c] = (is_spec ? vx->spec * s->params[3*col + c] : s->params[3*col + c] / PI) * g3;
            } else {
                g[c] = 0.0;
            }
        }
    }
}

/* Camera::sample, camera.hpp:51-60 */
static v3 camera_sample(const drtb_camera* c, int x, int y, stream_t* rng)
{
    double s = (x + uniform(rng)) / c->width;
    double t = (y + uniform(rng)) / c->height;
    double aspect = (double)c->width / c->height;
    v3 fw = {c->forward[0], c->forward[1], c->forward[2]};
    v3 rt = {c->right[0], c->right[1], c->right[2]};
    v3 up = {c->up[0], c->up[1], c->up[2]};
    v3 nup = {-1.0*up.x, -1.0*up.y, -1.0*up.z};
    v3 dir = fw;
    dir = add3(dir, mul3(rt, (2.*s - 1.) * aspect * tan(c->vfov / 2.)));
    dir = add3(dir, mul3(nup, (2.*t - 1.) * tan(c->vfov / 2.)));
    return normalize3(dir);
}

static int row_in_shard(int y, const drtb_render_opts* o)
{
    if (o->shard_count <= 1) return 1;
    int band = o->band_rows > 0 ? o->band_rows : 1;
    return (y / band) % o->shard_count == o->shard_index;
}

/* The pixel loop, src/render.cpp:72-86, gradient seed per SAMPLE, not divided
 * by spp or pdf.  Same contract as drtb_render. */
/* gimg != NULL: additionally the per-pixel gradient image of parameter gparam
 * (drtb_render_grad_image): each pixel's samples accumulate into a zeroed
 * scratch gradient that is then added to the totals. */
int drt_oracle_render_gimg(const drtb_scene* s, const drtb_mesh* mesh, const drtb_render_opts* o,
                           const double* seed_img, double* img, double* grad,
                           int gparam, double* gimg, int n_threads, drtb_stats* stats)
{
    if (mesh && mesh->n_triangles == 0) mesh = NULL;
    if (gimg && (gparam < 0 || gparam >= s->n_params)) return -1;
    const int W = s->camera.width, H = s->camera.height, spp = o->spp;
    const int P = s->n_params;
    int* rows = (int*)malloc(sizeof(int) * (size_t)(H > 0 ? H : 1));
    int nrows = 0;
    for (int y = 0; y < H; ++y) if (row_in_shard(y, o)) rows[nrows++] = y;
    if (n_threads < 1) n_threads = 1;
    double* gsum = (double*)calloc((size_t)n_threads * P * 3 + 1, sizeof(double));
    uint64_t segments = 0, lit = 0;
    const int want_grad = (o->flags & DRTB_FLAG_GRAD) != 0;
    const uint64_t key0 = o->seed * 0x9E3779B97F4A7C15ull;

#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads) reduction(+ : segments, lit)
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        path_buf buf; memset(&buf, 0, sizeof buf);
        counters_t cnt = {0, 0};
        double* g = gsum + (size_t)tid * P * 3;
        double* gpix = gimg ? (double*)calloc((size_t)P * 3 + 1, sizeof(double)) : NULL;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int r = 0; r < nrows; ++r) {
            int y = rows[r];
            for (int x = 0; x < W; ++x) {
                double seed[3] = {o->seed_scale, o->seed_scale, o->seed_scale};
                if (seed_img)
                    for (int c = 0; c < 3; ++c)
                        seed[c] = o->seed_scale * seed_img[((size_t)r * W + x) * 3 + c];
                double acc[3] = {0, 0, 0};
                for (int i = 0; i < spp; ++i) {
                    stream_t rng = {key0 + ((uint64_t)y * W + x) * spp + i, 0};
                    v3 eye = {s->camera.eye[0], s->camera.eye[1], s->camera.eye[2]};
                    v3 dir = camera_sample(&s->camera, x, y, &rng);
                    double L[3];
                    trace_path(s, mesh, o, &rng, eye, dir, &buf, L,
                               want_grad ? seed : NULL, gpix ? gpix : g, &cnt);
                    for (int c = 0; c < 3; ++c) acc[c] += L[c] / 1.0;   /* /pdf, pdf = 1 */
                }
                if (gpix) {
                    for (int c = 0; c < 3; ++c)
                        gimg[((size_t)r * W + x) * 3 + c] = gpix[3*gparam + c];
                    for (int j = 0; j < P * 3; ++j) { g[j] += gpix[j]; gpix[j] = 0.0; }
                }
                if (img)
                    for (int c = 0; c < 3; ++c)
                        img[((size_t)r * W + x) * 3 + c] = acc[c] / (double)spp;
            }
        }
        free(buf.v); free(buf.L); free(gpix);
        segments += cnt.segments; lit += cnt.lit;
    }
    if (grad)
        for (int j = 0; j < P * 3; ++j) {
            double a = 0;
            for (int t = 0; t < n_threads; ++t) a += gsum[(size_t)t * P * 3 + j];
            grad[j] = a;
        }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->paths = (uint64_t)nrows * W * spp;
        stats->segments = segments;
        stats->lit_paths = lit;
    }
    free(gsum); free(rows);
    return 0;
}

int drt_oracle_render_mesh(const drtb_scene* s, const drtb_mesh* mesh, const drtb_render_opts* o,
                           const double* seed_img, double* img, double* grad,
                           int n_threads, drtb_stats* stats)
{
    return drt_oracle_render_gimg(s, mesh, o, seed_img, img, grad, -1, NULL, n_threads, stats);
}

int drt_oracle_render(const drtb_scene* s, const drtb_render_opts* o,
                      const double* seed_img, double* img, double* grad,
                      int n_threads, drtb_stats* stats)
{
    return drt_oracle_render_mesh(s, NULL, o, seed_img, img, grad, n_threads, stats);
}

/* Pathtracer::trace on explicit rays (same contract as drtb_trace_rays). */
int drt_oracle_trace_rays_mesh(const drtb_scene* s, const drtb_mesh* mesh, const drtb_render_opts* o,
                               int64_t n, const double* orig, const double* dir,
                               const uint64_t* keys, double* radiance, double* jac)
{
    if (mesh && mesh->n_triangles == 0) mesh = NULL;
    const int P = s->n_params;
    path_buf buf; memset(&buf, 0, sizeof buf);
    counters_t cnt = {0, 0};
    double* g = (double*)calloc((size_t)P * 3 + 1, sizeof(double));
    for (int64_t i = 0; i < n; ++i) {
        v3 og = {orig[3*i], orig[3*i+1], orig[3*i+2]};
        v3 d = {dir[3*i], dir[3*i+1], dir[3*i+2]};
        double one[3] = {1, 1, 1};
        stream_t rng = {keys[i], 2};
        memset(g, 0, sizeof(double) * (size_t)P * 3);
        /* channels never mix, so one all-ones seed yields the whole diagonal */
        trace_path(s, mesh, o, &rng, og, d, &buf, radiance + 3*i, jac ? one : NULL, g, &cnt);
        if (jac) memcpy(jac + (size_t)i * P * 3, g, sizeof(double) * (size_t)P * 3);
    }
    free(g); free(buf.v); free(buf.L);
    return 0;
}

int drt_oracle_trace_rays(const drtb_scene* s, const drtb_render_opts* o,
                          int64_t n, const double* orig, const double* dir,
                          const uint64_t* keys, double* radiance, double* jac)
{
    return drt_oracle_trace_rays_mesh(s, NULL, o, n, orig, dir, keys, radiance, jac);
}

uint32_t drt_oracle_stream_draw(uint64_t key, uint32_t slot)
{
    return (uint32_t)(splitmix64(key * 0x100000001B3ull + slot) % 2147483647ull);
}

int drt_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
