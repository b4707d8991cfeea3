/*
 * vtrace_oracle.c — CPU ORACLE (test infrastructure only; see vtrace_oracle.h).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 * Every float operation is a separately rounded IEEE-754 binary32 op in the order
 * written; the CUDA path must reproduce the same decisions bit for bit.
 *
 * Reference anchors are cited as file:line relative to the reference tree.
 */
#include "vtrace_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* small vector / matrix helpers.  mat4 is column-major: c[col][row] (GLSL / glm-rs).        */

typedef struct { float c[4][4]; } mat4;
typedef struct { float v[4]; } vec4;

static mat4 mat4_load(const float* p) { mat4 m; memcpy(&m, p, sizeof m); return m; }

/* GLSL min(): IEEE-754 minNum semantics are FIXED here (the GLSL spec leaves NaN open). */
static inline float vo_fmin(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}

/* float -> int conversion with the saturating / NaN->0 behaviour of cvt.rzi.s32.f32 */
static inline int32_t vo_f2i(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

/* mat4 * vec4, GLSL OpMatrixTimesVector: sum over columns, left to right */
static inline vec4 mat4_mul_vec4(const mat4* m, vec4 a) {
    vec4 r;
    for (int i = 0; i < 4; ++i)
        r.v[i] = ((m->c[0][i] * a.v[0] + m->c[1][i] * a.v[1]) + m->c[2][i] * a.v[2]) + m->c[3][i] * a.v[3];
    return r;
}

/* mat4 * mat4, GLSL OpMatrixTimesMatrix: result column j = a * b[j] */
static mat4 mat4_mul(const mat4* a, const mat4* b) {
    mat4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r.c[j][i] = ((a->c[0][i] * b->c[j][0] + a->c[1][i] * b->c[j][1]) + a->c[2][i] * b->c[j][2]) +
                        a->c[3][i] * b->c[j][3];
    return r;
}

/* GLSL inverse(mat4): cofactor expansion over 2x2 sub-determinants, one reciprocal. */
static mat4 mat4_inverse(const mat4* m) {
#define A(r, c_) m->c[c_][r]
    float s0 = A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1);
    float s1 = A(0, 0) * A(1, 2) - A(1, 0) * A(0, 2);
    float s2 = A(0, 0) * A(1, 3) - A(1, 0) * A(0, 3);
    float s3 = A(0, 1) * A(1, 2) - A(1, 1) * A(0, 2);
    float s4 = A(0, 1) * A(1, 3) - A(1, 1) * A(0, 3);
    float s5 = A(0, 2) * A(1, 3) - A(1, 2) * A(0, 3);
    float c5 = A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3);
    float c4 = A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3);
    float c3 = A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2);
    float c2 = A(2, 0) * A(3, 3) - A(3, 0) * A(2, 3);
    float c1 = A(2, 0) * A(3, 2) - A(3, 0) * A(2, 2);
    float c0 = A(2, 0) * A(3, 1) - A(3, 0) * A(2, 1);
    float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    float id = 1.0f / det;
    mat4 r;
#define B(r_, c_) r.c[c_][r_]
    B(0, 0) = ((A(1, 1) * c5 - A(1, 2) * c4) + A(1, 3) * c3) * id;
    B(0, 1) = ((-A(0, 1) * c5 + A(0, 2) * c4) - A(0, 3) * c3) * id;
    B(0, 2) = ((A(3, 1) * s5 - A(3, 2) * s4) + A(3, 3) * s3) * id;
    B(0, 3) = ((-A(2, 1) * s5 + A(2, 2) * s4) - A(2, 3) * s3) * id;
    B(1, 0) = ((-A(1, 0) * c5 + A(1, 2) * c2) - A(1, 3) * c1) * id;
    B(1, 1) = ((A(0, 0) * c5 - A(0, 2) * c2) + A(0, 3) * c1) * id;
    B(1, 2) = ((-A(3, 0) * s5 + A(3, 2) * s2) - A(3, 3) * s1) * id;
    B(1, 3) = ((A(2, 0) * s5 - A(2, 2) * s2) + A(2, 3) * s1) * id;
    B(2, 0) = ((A(1, 0) * c4 - A(1, 1) * c2) + A(1, 3) * c0) * id;
    B(2, 1) = ((-A(0, 0) * c4 + A(0, 1) * c2) - A(0, 3) * c0) * id;
    B(2, 2) = ((A(3, 0) * s4 - A(3, 1) * s2) + A(3, 3) * s0) * id;
    B(2, 3) = ((-A(2, 0) * s4 + A(2, 1) * s2) - A(2, 3) * s0) * id;
    B(3, 0) = ((-A(1, 0) * c3 + A(1, 1) * c1) - A(1, 2) * c0) * id;
    B(3, 1) = ((A(0, 0) * c3 - A(0, 1) * c1) + A(0, 2) * c0) * id;
    B(3, 2) = ((-A(3, 0) * s3 + A(3, 1) * s1) - A(3, 2) * s0) * id;
    B(3, 3) = ((A(2, 0) * s3 - A(2, 1) * s1) + A(2, 2) * s0) * id;
#undef A
#undef B
    return r;
}

void vo_mat4_inverse(const float* m, float* out) {
    mat4 a = mat4_load(m), r = mat4_inverse(&a);
    memcpy(out, &r, sizeof r);
}
void vo_mat4_mul(const float* a, const float* b, float* out) {
    mat4 x = mat4_load(a), y = mat4_load(b), r = mat4_mul(&x, &y);
    memcpy(out, &r, sizeof r);
}

/* ------------------------------------------------------------------------------------------ */
/* sRGB: VK_FORMAT_R8G8B8A8_SRGB textures (lib/memory.c:317) and the B8G8R8A8_SRGB target    */
/* (lib/swapchain.c:88).  Decode = 256-entry table; encode = threshold table so that          */
/* encode(decode(c)) == c and both are exact table look-ups (no pow on the hot path).         */

static float g_srgb_dec[256];
static float g_srgb_thr[256]; /* thr[k] = linear value of sRGB code (k - 0.5)/255, k >= 1 */
static int g_srgb_ready = 0;

static double srgb_to_linear_d(double c) { return c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4); }

static void srgb_init(void) {
    if (g_srgb_ready) return;
    for (int k = 0; k < 256; ++k) {
        g_srgb_dec[k] = (float)srgb_to_linear_d((double)k / 255.0);
        g_srgb_thr[k] = k == 0 ? 0.0f : (float)srgb_to_linear_d(((double)k - 0.5) / 255.0);
    }
    g_srgb_ready = 1;
}

float vo_srgb_decode(uint8_t c) { srgb_init(); return g_srgb_dec[c]; }

uint8_t vo_srgb_encode(float x) {
    srgb_init();
    /* largest k with x >= thr[k]; NaN compares false everywhere -> 0 */
    uint32_t k = 0;
    for (uint32_t bit = 128; bit; bit >>= 1)
        if (x >= g_srgb_thr[k | bit]) k |= bit;
    return (uint8_t)k;
}

/* ------------------------------------------------------------------------------------------ */
/* scene state = what crosses the C ABI                                       */

typedef struct {
    uint32_t w, h, d; uint8_t* rgba; uint32_t kind, seed; float* heights;
    uint32_t* btable; uint32_t* bmasks; uint8_t* bcolors; /* VO_VOLUME_BRICKS: slot per brick (or ~0), 16 words + RGBA per slot */
} vo_texture;

/* what the traversal needs to know about a volume.  kind 0 = dense RGBA8 texels (add_texture);
 * kinds 1/2 = procedural volumes of the large-scene extension (no reference counterpart). */
typedef struct {
    uint32_t kind, w, h, d, seed; const uint8_t* rgba; const float* heights;
    const uint32_t* btable; const uint32_t* bmasks; const uint8_t* bcolors;
} vol_view;

static inline uint32_t vo_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static inline uint32_t vo_hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
    uint32_t h = vo_mix32(x * 0x9E3779B1u + seed);
    h = vo_mix32(h ^ (y * 0x85EBCA77u));
    h = vo_mix32(h ^ (z * 0xC2B2AE3Du));
    return h;
}

/* VO_VOLUME_HEIGHTMAP: height(x,z) = H/2 + H/4 * fbm, fbm = 5 octaves of value noise on lattices of
 * 128, 64, 32, 16, 8 voxels with amplitudes 1/2 .. 1/32; a voxel is filled iff its altitude
 * (H - 1 - y, because +Y is down) satisfies (float)altitude <= height. */
static float heightmap_height(uint32_t x, uint32_t z, uint32_t H, uint32_t seed) {
    float n = 0.0f, amp = 0.5f;
    for (uint32_t o = 0; o < 5; ++o) {
        const uint32_t cell = 128u >> o;
        const uint32_t ix = x / cell, iz = z / cell;
        const float fx = (float)(x % cell) / (float)cell, fz = (float)(z % cell) / (float)cell;
        const float ux = (fx * fx) * (3.0f - 2.0f * fx), uz = (fz * fz) * (3.0f - 2.0f * fz);
        const float v00 = (float)(vo_hash3(ix, iz, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float v10 = (float)(vo_hash3(ix + 1, iz, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float v01 = (float)(vo_hash3(ix, iz + 1, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float v11 = (float)(vo_hash3(ix + 1, iz + 1, o, seed) >> 8) * (1.0f / 16777216.0f);
        const float a = v00 + ux * (v10 - v00);
        const float b = v01 + ux * (v11 - v01);
        const float v = a + uz * (b - a);
        n = n + amp * (2.0f * v - 1.0f);
        amp = amp * 0.5f;
    }
    return 0.5f * (float)H + (0.25f * (float)H) * n;
}

/* texel (x,y,z) of a volume -> RGBA8; alpha 0 = empty */
static inline void vol_texel(const vol_view* v, int32_t x, int32_t y, int32_t z, uint8_t out[4]) {
    if (v->kind == 0) {
        memcpy(out, v->rgba + 4 * ((size_t)x + (size_t)v->w * ((size_t)y + (size_t)v->h * (size_t)z)), 4);
    } else if (v->kind == VO_VOLUME_HEIGHTMAP) {
        const float hgt = v->heights[(size_t)z * v->w + (size_t)x];
        /* world +Y points DOWN on screen in the reference, so the ground fills the high-y side */
        const uint32_t alt = v->h - 1u - (uint32_t)y;
        if ((float)alt <= hgt) {
            const uint32_t band = (alt * 4u) / v->h; /* colour by altitude band */
            static const uint8_t pal[4][3] = {{72, 60, 50}, {96, 128, 56}, {120, 120, 120}, {240, 240, 245}};
            out[0] = pal[band][0]; out[1] = pal[band][1]; out[2] = pal[band][2]; out[3] = 255;
        } else {
            out[0] = out[1] = out[2] = out[3] = 0;
        }
    } else if (v->kind == VO_VOLUME_BRICKS) { /* caller-supplied 8^3 bricks: occupancy bits + one colour per brick */
        const uint32_t ux = (uint32_t)x, uy = (uint32_t)y, uz = (uint32_t)z;
        const size_t b = ((size_t)(uz >> 3) * (v->h >> 3) + (uy >> 3)) * (v->w >> 3) + (ux >> 3);
        const uint32_t slot = v->btable[b];
        out[0] = out[1] = out[2] = out[3] = 0;
        if (slot != 0xFFFFFFFFu) {
            const uint32_t wv = v->bmasks[(size_t)slot * 16 + (((uz & 7u) << 1) | ((uy & 7u) >> 2))];
            if ((wv >> ((ux & 7u) | ((uy & 3u) << 3))) & 1u) {
                memcpy(out, v->bcolors + 4 * (size_t)slot, 3);
                out[3] = 255;
            }
        }
    } else { /* VO_VOLUME_SPARSE_BRICKS: 2 % of the 8^3 bricks are non-empty, half of their voxels filled */
        const uint32_t ux = (uint32_t)x, uy = (uint32_t)y, uz = (uint32_t)z;
        const int brick = (vo_hash3(ux >> 3, uy >> 3, uz >> 3, v->seed) & 0xFFFFu) < 1311u;
        if (brick && (vo_hash3(ux, uy, uz, v->seed ^ 0x5bd1e995u) & 1u)) {
            const uint32_t c = vo_hash3(ux, uy, uz, v->seed ^ 0x27d4eb2fu);
            out[0] = (uint8_t)(c | 0x40u); out[1] = (uint8_t)((c >> 8) | 0x40u); out[2] = (uint8_t)((c >> 16) | 0x40u); out[3] = 255;
        } else {
            out[0] = out[1] = out[2] = out[3] = 0;
        }
    }
}

struct vo_scene {
    vo_texture* tex;
    uint32_t ntex, captex;
    float* inst;
    uint32_t ninst, capinst;
};

vo_scene* vo_scene_create(void) {
    vo_scene* s = (vo_scene*)calloc(1, sizeof *s);
    s->capinst = 1;
    s->inst = (float*)calloc(16, sizeof(float)); /* zero-filled "stale" instance 0 */
    s->ninst = 1;                                /* lib/memory.c:236,251: 0 is coerced to 1 */
    return s;
}

void vo_scene_destroy(vo_scene* s) {
    if (!s) return;
    for (uint32_t i = 0; i < s->ntex; ++i) {
        free(s->tex[i].rgba); free(s->tex[i].heights); free(s->tex[i].btable); free(s->tex[i].bmasks); free(s->tex[i].bcolors);
    }
    free(s->tex);
    free(s->inst);
    free(s);
}

int32_t vo_add_texture(vo_scene* s, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t d) {
    if (s->ntex >= 65536u) return -1; /* MAX_TEXTURES, lib/common.h:35, lib/memory.c:287-290 */
    if (s->ntex == s->captex) {
        s->captex = s->captex ? s->captex * 2 : 4;
        s->tex = (vo_texture*)realloc(s->tex, s->captex * sizeof *s->tex);
    }
    size_t bytes = (size_t)4 * w * h * d;
    vo_texture* t = &s->tex[s->ntex];
    t->w = w; t->h = h; t->d = d;
    t->kind = 0; t->seed = 0; t->heights = NULL; t->btable = NULL; t->bmasks = NULL; t->bcolors = NULL;
    t->rgba = (uint8_t*)malloc(bytes ? bytes : 1);
    memcpy(t->rgba, rgba, bytes); /* lib/memory.c:304-307: data only borrowed for the call */
    return (int32_t)s->ntex++;
}

int32_t vo_add_volume_procedural(vo_scene* s, uint32_t kind, uint32_t w, uint32_t h, uint32_t d, uint32_t seed) {
    if (s->ntex >= 65536u || (kind != VO_VOLUME_HEIGHTMAP && kind != VO_VOLUME_SPARSE_BRICKS)) return -1;
    if (s->ntex == s->captex) {
        s->captex = s->captex ? s->captex * 2 : 4;
        s->tex = (vo_texture*)realloc(s->tex, s->captex * sizeof *s->tex);
    }
    vo_texture* t = &s->tex[s->ntex];
    t->w = w; t->h = h; t->d = d; t->kind = kind; t->seed = seed; t->rgba = NULL; t->heights = NULL;
    t->btable = NULL; t->bmasks = NULL; t->bcolors = NULL;
    if (kind == VO_VOLUME_HEIGHTMAP) {
        t->heights = (float*)malloc(sizeof(float) * (size_t)w * d);
#ifdef _OPENMP
#pragma omp parallel for
#endif
        for (int64_t z = 0; z < (int64_t)d; ++z)
            for (uint32_t x = 0; x < w; ++x) t->heights[(size_t)z * w + x] = heightmap_height(x, (uint32_t)z, h, seed);
    }
    return (int32_t)s->ntex++;
}

int32_t vo_add_volume_bricks(vo_scene* s, const uint32_t* coords, const uint32_t* masks, const uint8_t* colors, uint64_t n,
                             uint32_t w, uint32_t h, uint32_t d) {
    if (s->ntex >= 65536u || !w || !h || !d || ((w | h | d) & 7u)) return -1;
    if (s->ntex == s->captex) {
        s->captex = s->captex ? s->captex * 2 : 4;
        s->tex = (vo_texture*)realloc(s->tex, s->captex * sizeof *s->tex);
    }
    vo_texture* t = &s->tex[s->ntex];
    memset(t, 0, sizeof *t);
    t->w = w; t->h = h; t->d = d; t->kind = VO_VOLUME_BRICKS;
    const size_t nb = (size_t)(w >> 3) * (h >> 3) * (d >> 3);
    t->btable = (uint32_t*)malloc(nb * 4);
    memset(t->btable, 0xFF, nb * 4);
    t->bmasks = (uint32_t*)malloc((size_t)(n ? n : 1) * 64);
    t->bcolors = (uint8_t*)malloc((size_t)(n ? n : 1) * 4);
    memcpy(t->bmasks, masks, (size_t)n * 64);
    memcpy(t->bcolors, colors, (size_t)n * 4);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t bx = coords[3 * i], by = coords[3 * i + 1], bz = coords[3 * i + 2];
        if (bx >= (w >> 3) || by >= (h >> 3) || bz >= (d >> 3)) return -1;
        t->btable[((size_t)bz * (h >> 3) + by) * (w >> 3) + bx] = (uint32_t)i;
    }
    return (int32_t)s->ntex++;
}

void vo_set_instances(vo_scene* s, const float* mats, uint32_t n) {
    uint32_t n1 = n ? n : 1; /* lib/memory.c:236,251 */
    if (n1 > s->capinst) {
        s->inst = (float*)realloc(s->inst, (size_t)n1 * 16 * sizeof(float));
        memset(s->inst + (size_t)s->capinst * 16, 0, (size_t)(n1 - s->capinst) * 16 * sizeof(float));
        s->capinst = n1;
    }
    if (n) memcpy(s->inst, mats, (size_t)n * 16 * sizeof(float));
    s->ninst = n1;
}

/* ------------------------------------------------------------------------------------------ */
/* uniforms                                                                                   */

typedef struct {
    mat4 P, V, Pi, Vi, Vci, RD, PV;
    float eye[3];
    float vw, vh;
    float sxn, syn; /* 2/vw, 2/vh: pixel -> NDC scale */
    float sun[3];   /* unit vector towards the sun, world space (shadow-ray extension) */
    int width, height;
} frame_uniforms;

typedef struct {
    mat4 M, Mi, MVP;
    float dirm[4][3]; /* Mi(3x3) * RD(rows 0-2): clip-space point -> model-space ray direction */
    float eye_m[3];
    float sun_m[3]; /* inverse(M)3x3 * sun: direction of the shadow rays in model space */
    uint32_t tex;
    uint32_t w, h, d;
    vol_view vol;
    int valid;
} inst_uniforms;

static void frame_setup(frame_uniforms* F, const float* P, const float* V, int width, int height, uint32_t flags) {
    F->P = mat4_load(P);
    F->V = mat4_load(V);
    F->Pi = mat4_inverse(&F->P);              /* trace.frag:48 */
    F->Vi = mat4_inverse(&F->V);              /* trace.frag:49 */
    mat4 Vc = F->V;                           /* trace.frag:54-57 */
    Vc.c[3][0] = 0.0f; Vc.c[3][1] = 0.0f; Vc.c[3][2] = 0.0f;
    F->Vci = mat4_inverse(&Vc);               /* trace.frag:59 */
    F->RD = mat4_mul(&F->Vci, &F->Pi);        /* trace.frag:59: (inverse(Vc) * inverse(P)) * sp */
    F->PV = mat4_mul(&F->P, &F->V);           /* trace.vert:45: (P * V) * world_position */
    vec4 o = {{0.0f, 0.0f, 0.0f, 1.0f}};
    vec4 e = mat4_mul_vec4(&F->Vi, o);        /* trace.frag:51 cam_pos */
    F->eye[0] = e.v[0]; F->eye[1] = e.v[1]; F->eye[2] = e.v[2];
    F->width = width; F->height = height;
    F->vw = (float)width;                     /* lib/command.c:80 */
    F->vh = (flags & VO_FLAG_VIEWPORT_H_IS_W) ? (float)width : (float)height; /* lib/command.c:81 */
    F->sxn = 2.0f / F->vw;
    F->syn = 2.0f / F->vh;
    {   /* sun direction (0.4, -0.8, 0.45), normalised */
        const float sx = 0.4f, sy = -0.8f, sz = 0.45f;
        const float l = sqrtf((sx * sx + sy * sy) + sz * sz);
        F->sun[0] = sx / l; F->sun[1] = sy / l; F->sun[2] = sz / l;
    }
}

static void inst_setup(inst_uniforms* I, const frame_uniforms* F, const vo_scene* s, const float* m16) {
    I->M = mat4_load(m16);
    memcpy(&I->tex, &I->M.c[3][3], 4);        /* trace.vert:38 floatBitsToInt(model[3][3]) */
    I->M.c[3][3] = 1.0f;                      /* trace.vert:39-40 */
    I->valid = I->tex < s->ntex;
    if (!I->valid) return;
    I->w = s->tex[I->tex].w; I->h = s->tex[I->tex].h; I->d = s->tex[I->tex].d;
    I->vol.kind = s->tex[I->tex].kind; I->vol.w = I->w; I->vol.h = I->h; I->vol.d = I->d;
    I->vol.seed = s->tex[I->tex].seed; I->vol.rgba = s->tex[I->tex].rgba; I->vol.heights = s->tex[I->tex].heights;
    I->vol.btable = s->tex[I->tex].btable; I->vol.bmasks = s->tex[I->tex].bmasks; I->vol.bcolors = s->tex[I->tex].bcolors;
    I->Mi = mat4_inverse(&I->M);              /* trace.frag:65 */
    I->MVP = mat4_mul(&F->PV, &I->M);
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 3; ++i)
            I->dirm[j][i] = (I->Mi.c[0][i] * F->RD.c[j][0] + I->Mi.c[1][i] * F->RD.c[j][1]) + I->Mi.c[2][i] * F->RD.c[j][2];
    for (int i = 0; i < 3; ++i)
        I->sun_m[i] = (I->Mi.c[0][i] * F->sun[0] + I->Mi.c[1][i] * F->sun[1]) + I->Mi.c[2][i] * F->sun[2];
    vec4 e = {{F->eye[0], F->eye[1], F->eye[2], 1.0f}};
    vec4 em = mat4_mul_vec4(&I->Mi, e);
    I->eye_m[0] = em.v[0]; I->eye_m[1] = em.v[1]; I->eye_m[2] = em.v[2];
}

/* ------------------------------------------------------------------------------------------ */
/* ray / unit-cube slab test (the rasteriser restatement: which point of the proxy cube's     */
/* FRONT faces covers the sample; lib/memory.c:22-40 cube, lib/pipeline.c:120-121 culling).   */
/* Returns 1 and (tn, axis) when the ray o + t d, t > 0, enters [-0.5,0.5]^3 from outside.    */

static int slab_unit_cube(const float o[3], const float d[3], float* tn_out, int* axis_out) {
    float tn = -INFINITY, tf = INFINITY;
    int axis = -1;
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0f) {
            if (o[k] < -0.5f || o[k] > 0.5f) return 0;
            continue;
        }
        float inv = 1.0f / d[k];
        float t1 = (-0.5f - o[k]) * inv;
        float t2 = (0.5f - o[k]) * inv;
        float lo = t1 < t2 ? t1 : t2;
        float hi = t1 < t2 ? t2 : t1;
        if (lo > tn) { tn = lo; axis = k; }
        if (hi < tf) tf = hi;
    }
    if (axis < 0) return 0;
    if (!(tn <= tf)) return 0;
    if (!(tn > 0.0f)) return 0; /* origin inside / box behind: only back faces visible -> culled */
    *tn_out = tn;
    *axis_out = axis;
    return 1;
}

static void entry_point(const float o[3], const float d[3], float tn, int axis, float mp[3]) {
    for (int k = 0; k < 3; ++k) {
        float p = o[k] + tn * d[k];
        p = p < -0.5f ? -0.5f : p;
        p = p > 0.5f ? 0.5f : p;
        mp[k] = p;
    }
    mp[axis] = d[axis] > 0.0f ? -0.5f : 0.5f;
}

/* ------------------------------------------------------------------------------------------ */
/* the DDA: shaders/trace.frag:63-89, verbatim                                                */

typedef struct {
    int hit;
    int32_t voxel[3];
    uint32_t steps;
    uint32_t last_mask; /* bit k: axis k advanced in the last executed iteration */
    int32_t step[3];
    float side[3], delta[3];
    float dir[3], len;
    float pos[3];
    uint8_t rgba[4];
} dda_state;

/* texel fetched by texture(tex, voxel / size) with a NEAREST, unnormalised-lower-edge coordinate
 * (trace.frag:76, lib/descriptor.c:100-115): i = clamp(floor(fl(v / s) * s), 0, s - 1).       */
static inline int32_t texel_of(int32_t v, float size, int32_t isize) {
    float u = (float)v / size;
    int32_t i = vo_f2i(floorf(u * size));
    if (i < 0) i = 0;
    if (i > isize - 1) i = isize - 1;
    return i;
}

static void dda_march(const vol_view* vol, const float pos[3], const float dir[3], const int32_t* start_voxel,
                      dda_state* r) {
    const uint32_t W = vol->w, H = vol->h, D = vol->d;
    const int32_t isz[3] = {(int32_t)W, (int32_t)H, (int32_t)D};
    const float size[3] = {(float)isz[0], (float)isz[1], (float)isz[2]}; /* :63-64 */
    float sgn[3];
    r->len = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]); /* length(), :70 */
    for (int k = 0; k < 3; ++k) {
        r->pos[k] = pos[k];
        r->dir[k] = dir[k];
        /* :68 ivec3(floor(min(pos, size - 1))) */
        r->voxel[k] = start_voxel ? start_voxel[k] : vo_f2i(floorf(vo_fmin(pos[k], size[k] - 1.0f)));
        sgn[k] = dir[k] > 0.0f ? 1.0f : (dir[k] < 0.0f ? -1.0f : 0.0f); /* sign(), :69 */
        r->step[k] = (int32_t)sgn[k];
        r->delta[k] = fabsf(r->len / dir[k]);                           /* :70 */
        r->side[k] = ((sgn[k] * ((float)r->voxel[k] - pos[k]) + sgn[k] * 0.5f) + 0.5f) * r->delta[k]; /* :71 */
    }
    r->steps = 0;
    r->last_mask = 0;
    r->hit = 0;
    const uint32_t max_steps = W + H + D; /* :74 */
    while (r->steps < max_steps && r->voxel[0] >= 0 && r->voxel[1] >= 0 && r->voxel[2] >= 0 &&
           r->voxel[0] < isz[0] && r->voxel[1] < isz[1] && r->voxel[2] < isz[2]) { /* :75 */
        int32_t tx = texel_of(r->voxel[0], size[0], isz[0]);
        int32_t ty = texel_of(r->voxel[1], size[1], isz[1]);
        int32_t tz = texel_of(r->voxel[2], size[2], isz[2]);
        uint8_t s[4];
        vol_texel(vol, tx, ty, tz, s); /* :76 */
        if (s[3] > 0) { /* :78 texSample.w > 0.0 */
            memcpy(r->rgba, s, 4);
            r->hit = 1;
            return; /* :79-80 */
        }
        /* :83 mask = lessThanEqual(side.xyz, min(side.yzx, side.zxy)) */
        int m0 = r->side[0] <= vo_fmin(r->side[1], r->side[2]);
        int m1 = r->side[1] <= vo_fmin(r->side[2], r->side[0]);
        int m2 = r->side[2] <= vo_fmin(r->side[0], r->side[1]);
        /* :84 side += vec3(mask) * delta  (0 * inf = NaN is kept on purpose) */
        r->side[0] += (m0 ? 1.0f : 0.0f) * r->delta[0];
        r->side[1] += (m1 ? 1.0f : 0.0f) * r->delta[1];
        r->side[2] += (m2 ? 1.0f : 0.0f) * r->delta[2];
        /* :85 voxel += ivec3(mask) * step */
        r->voxel[0] += m0 * r->step[0];
        r->voxel[1] += m1 * r->step[1];
        r->voxel[2] += m2 * r->step[2];
        r->last_mask = (uint32_t)(m0 | (m1 << 1) | (m2 << 2));
        ++r->steps; /* :86 */
    }
    /* :89 discard */
}

/* fragment-stage prologue of trace.frag: :59 ray_dir, :65 model_ray_dir, :66 model_ray_pos */
static void frag_ray(const mat4* RD, const mat4* Mi, vec4 sp, const float mp[3], uint32_t W, uint32_t H,
                     uint32_t D, float pos[3], float dir[3]) {
    vec4 r = mat4_mul_vec4(RD, sp);
    float len = sqrtf((r.v[0] * r.v[0] + r.v[1] * r.v[1]) + r.v[2] * r.v[2]);
    vec4 rd = {{r.v[0] / len, r.v[1] / len, r.v[2] / len, 0.0f}}; /* normalize(), :59 */
    vec4 md = mat4_mul_vec4(Mi, rd);                              /* :65 */
    const float size[3] = {(float)(int32_t)W, (float)(int32_t)H, (float)(int32_t)D};
    for (int k = 0; k < 3; ++k) {
        dir[k] = md.v[k];
        pos[k] = (mp[k] + 0.5f) * size[k]; /* :66 */
    }
}

void vo_frag_main(const float* P, const float* V, const float* M, const float* screen_position,
                  const float* model_position, const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t d,
                  int32_t* out, float* color, float* frag_depth) {
    srgb_init();
    frame_uniforms F;
    frame_setup(&F, P, V, 1, 1, 0);
    mat4 Mm = mat4_load(M);
    mat4 Mi = mat4_inverse(&Mm);
    vec4 sp = {{screen_position[0], screen_position[1], screen_position[2], screen_position[3]}};
    if (frag_depth) *frag_depth = sp.v[2] / sp.v[3]; /* :46 */
    float pos[3], dir[3];
    frag_ray(&F.RD, &Mi, sp, model_position, w, h, d, pos, dir);
    dda_state r;
    const vol_view vv = {0, w, h, d, 0, rgba, NULL, NULL, NULL, NULL};
    dda_march(&vv, pos, dir, NULL, &r);
    out[0] = r.hit;
    out[1] = r.voxel[0]; out[2] = r.voxel[1]; out[3] = r.voxel[2];
    out[4] = (int32_t)r.steps;
    out[5] = (int32_t)r.last_mask;
    if (color) {
        if (r.hit) {
            color[0] = g_srgb_dec[r.rgba[0]]; color[1] = g_srgb_dec[r.rgba[1]];
            color[2] = g_srgb_dec[r.rgba[2]]; color[3] = (float)r.rgba[3] / 255.0f;
        } else {
            color[0] = color[1] = color[2] = color[3] = 0.0f;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* one fragment = rasteriser restatement + trace.frag                                          */

typedef struct {
    int covered;
    float depth;    /* gl_FragDepth, trace.frag:46 */
    int entry_axis;
    float mp[3];
    dda_state dda;
} fragment;

static void run_fragment(const frame_uniforms* F, const inst_uniforms* I, float fx, float fy, fragment* f) {
    f->covered = 0;
    f->dda.hit = 0;
    f->dda.steps = 0;
    if (!I->valid) return;
    /* NDC of the sample point, no Y flip anywhere */
    float x_ndc = fx * F->sxn - 1.0f;
    float y_ndc = fy * F->syn - 1.0f;
    float d[3];
    for (int k = 0; k < 3; ++k) d[k] = (I->dirm[0][k] * x_ndc + I->dirm[1][k] * y_ndc) + I->dirm[3][k];
    float tn;
    int axis;
    if (!slab_unit_cube(I->eye_m, d, &tn, &axis)) return;
    entry_point(I->eye_m, d, tn, axis, f->mp);
    /* trace.vert:43-45 at the covered point: screen_position = (P V M) * (mp, 1) */
    vec4 sp;
    for (int i = 0; i < 4; ++i)
        sp.v[i] = ((I->MVP.c[0][i] * f->mp[0] + I->MVP.c[1][i] * f->mp[1]) + I->MVP.c[2][i] * f->mp[2]) + I->MVP.c[3][i];
    /* Vulkan clip volume 0 <= z <= w (GL-style P used unmodified) */
    if (!(sp.v[3] > 0.0f && sp.v[2] >= 0.0f && sp.v[2] <= sp.v[3])) return;
    f->covered = 1;
    f->entry_axis = axis;
    f->depth = sp.v[2] / sp.v[3]; /* :46 */
    float pos[3], dir[3];
    frag_ray(&F->RD, &I->Mi, sp, f->mp, I->w, I->h, I->d, pos, dir);
    dda_march(&I->vol, pos, dir, NULL, &f->dda);
}

/* Where a secondary ray (bounce or shadow) leaves a hit: axis = first axis advanced by the last DDA
 * iteration (or the box-entry axis when steps == 0), origin = hit point clamped to the hit voxel with
 * the normal component on the face plane, start voxel = the neighbour across that face. */
static void leave_hit(const dda_state* r, int entry_axis, const float size[3], int* a_out, int* nsign_out, float p0[3],
                      int32_t sv[3]) {
    uint32_t lm = r->steps ? r->last_mask : (1u << entry_axis);
    int a = (lm & 1u) ? 0 : ((lm & 2u) ? 1 : 2);
    float t = r->steps ? r->side[a] - r->delta[a] : 0.0f;
    int nsign = r->step[a] != 0 ? -r->step[a] : (r->pos[a] <= 0.5f * size[a] ? -1 : 1);
    float tl = t / r->len;
    for (int k = 0; k < 3; ++k) {
        float p = r->pos[k] + r->dir[k] * tl;
        float lo = (float)r->voxel[k], hi = (float)(r->voxel[k] + 1);
        p = p < lo ? lo : p;
        p = p > hi ? hi : p;
        p0[k] = p;
        sv[k] = r->voxel[k];
    }
    p0[a] = (float)(r->voxel[a] + (nsign > 0 ? 1 : 0));
    sv[a] += nsign;
    *a_out = a;
    *nsign_out = nsign;
}

static inline uint32_t face_bits(const dda_state* r, int entry_axis) {
    uint32_t mask = r->steps ? r->last_mask : (1u << entry_axis);
    uint32_t neg = (uint32_t)(r->step[0] < 0) | ((uint32_t)(r->step[1] < 0) << 1) | ((uint32_t)(r->step[2] < 0) << 2);
    return mask | (neg << 3);
}

static uint64_t g_last_shadow_rays = 0;

uint64_t vo_render_primary(const vo_scene* s, const float* P, const float* V, int width, int height,
                           uint32_t flags, vo_hit_record* records, uint8_t* rgba8, float* depth_out,
                           int num_threads) {
    srgb_init();
    frame_uniforms F;
    frame_setup(&F, P, V, width, height, flags);
    inst_uniforms* I = (inst_uniforms*)malloc(sizeof(inst_uniforms) * s->ninst);
    for (uint32_t i = 0; i < s->ninst; ++i) inst_setup(&I[i], &F, s, s->inst + 16 * (size_t)i);
    /* clear values, lib/command.c:56-61, stored through the sRGB target */
    const uint8_t clear[4] = {vo_srgb_encode(53.0f / 100.0f), vo_srgb_encode(81.0f / 100.0f),
                              vo_srgb_encode(92.0f / 100.0f), 255};
    uint64_t total_iters = 0, shadow_rays = 0;
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads) reduction(+ : total_iters, shadow_rays)
#endif
    for (int py = 0; py < height; ++py) {
        for (int px = 0; px < width; ++px) {
            if ((flags & VO_FLAG_SUBSET_8) && ((px | py) & 7)) continue; /* 1/64 of the pixels (full-size parity checks) */
            float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
            uint8_t dst[4] = {clear[0], clear[1], clear[2], clear[3]};
            float zbuf = 1.0f; /* lib/command.c:60 */
            vo_hit_record rec = {VO_MISS, 0, VO_MISS, 0};
            for (uint32_t i = 0; i < s->ninst; ++i) { /* draw order = instance order, lib/command.c:102 */
                fragment f;
                run_fragment(&F, &I[i], fx, fy, &f);
                if (!f.covered) continue;
                rec.iters += f.dda.steps;
                if (!f.dda.hit) continue;                 /* discard, trace.frag:89 */
                if (!(f.depth < zbuf)) continue;          /* VK_COMPARE_OP_LESS, lib/pipeline.c:148-150 */
                zbuf = f.depth;
                /* extension: one shadow ray towards the sun, inside the fragment's own volume */
                float shade = 1.0f;
                uint32_t shadow_bits = 0;
                if (flags & VO_FLAG_SHADOW_RAYS) {
                    const float size[3] = {(float)(int32_t)I[i].w, (float)(int32_t)I[i].h, (float)(int32_t)I[i].d};
                    int ax, nsign;
                    float p0[3];
                    int32_t sv[3];
                    leave_hit(&f.dda, f.entry_axis, size, &ax, &nsign, p0, sv);
                    int lit = 0;
                    if ((float)nsign * I[i].sun_m[ax] > 0.0f) { /* the face looks at the sun */
                        dda_state sh;
                        dda_march(&I[i].vol, p0, I[i].sun_m, sv, &sh);
                        rec.iters += sh.steps;
                        lit = !sh.hit;
                        shadow_bits = 1u | ((uint32_t)lit << 1);
                        ++shadow_rays;
                    }
                    shade = lit ? 1.0f : 0.35f;
                }
                /* blend, lib/pipeline.c:129-137 */
                float a = (float)f.dda.rgba[3] / 255.0f;
                for (int c = 0; c < 3; ++c) {
                    float src = g_srgb_dec[f.dda.rgba[c]] * shade;
                    float dl = g_srgb_dec[dst[c]];
                    dst[c] = vo_srgb_encode(src * a + dl * (1.0f - a));
                }
                dst[3] = (uint8_t)vo_f2i(floorf(a * 255.0f + 0.5f));
                rec.hit_voxel = (uint32_t)f.dda.voxel[0] + I[i].w * ((uint32_t)f.dda.voxel[1] + I[i].h * (uint32_t)f.dda.voxel[2]);
                rec.packed = (f.dda.steps & 0xFFFFu) | (face_bits(&f.dda, f.entry_axis) << 16) | (shadow_bits << 22);
                rec.instance = i;
            }
            size_t p = (size_t)py * (size_t)width + (size_t)px;
            if (records) records[p] = rec;
            if (rgba8) memcpy(rgba8 + 4 * p, dst, 4);
            if (depth_out) depth_out[p] = zbuf;
            total_iters += rec.iters;
        }
    }
    free(I);
    g_last_shadow_rays = shadow_rays;
    return total_iters;
}

uint64_t vo_last_shadow_rays(void) { return g_last_shadow_rays; }

/* ------------------------------------------------------------------------------------------ */
/* path-tracing extension (no reference counterpart)                            */

static inline uint32_t vo_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
typedef struct { uint32_t key, ctr; } vo_rng;
/* counter-based stream keyed by (seed, pixel, sample): draw k is mix(key + k * phi) */
static inline void rng_init(vo_rng* r, uint32_t seed, uint32_t pixel, uint32_t sample) {
    uint32_t k = vo_mix(seed + pixel * 0x9E3779B9u);
    r->key = vo_mix(k ^ (sample * 0x85EBCA6Bu + 0xC2B2AE35u));
    r->ctr = 0;
}
/* uniform in [0,1) with 23 random mantissa bits: as_float(0x3f800000 | bits) - 1 (no int->float conversion) */
static inline float rng_u01(vo_rng* r) {
    uint32_t x = vo_mix(r->key + (r->ctr++) * 0x9E3779B9u);
    uint32_t u = 0x3f800000u | (x >> 9);
    float f;
    memcpy(&f, &u, 4);
    return f - 1.0f;
}

/* uniform point on the unit sphere, Marsaglia (1972): only + - * sqrt, so it is bit-reproducible */
static void rng_sphere(vo_rng* r, float s[3]) {
    float a = 0.0f, b = 0.0f, q = 0.0f;
    int ok = 0;
    for (int attempt = 0; attempt < 16 && !ok; ++attempt) {
        a = rng_u01(r) * 2.0f - 1.0f;
        b = rng_u01(r) * 2.0f - 1.0f;
        q = a * a + b * b;
        ok = q < 1.0f;
    }
    if (!ok) { a = 0.0f; b = 0.0f; q = 0.0f; }
    float w = sqrtf(1.0f - q);
    s[0] = (2.0f * a) * w;
    s[1] = (2.0f * b) * w;
    s[2] = 1.0f - 2.0f * q;
}

typedef struct {
    int hit;
    uint32_t instance;
    int entry_axis;
    dda_state dda;
} path_hit;

/* nearest instance along a ray, by box-entry parameter (ties: lower index); the first instance in
 * that order whose DDA hits wins.  `skip` is excluded.  Two kinds of ray:
 *   camera (cam != NULL): per instance o = eye in model space, d = dirm * (x_ndc, y_ndc, 1) — the
 *     same model-space ray the rasteriser restatement uses; cam = {x_ndc, y_ndc};
 *   world  (cam == NULL): o = Mi * (ow,1), d = Mi * (dw,0). */
static void trace_world(const inst_uniforms* I, uint32_t ninst, uint32_t skip, const float* cam, const float ow[3],
                        const float dw[3], path_hit* out, uint64_t* iters) {
    out->hit = 0;
    float last_t = -INFINITY;
    uint32_t last_j = 0;
    int have_last = 0;
    for (;;) {
        int found = 0;
        float best_t = 0.0f;
        uint32_t best_j = 0;
        int best_axis = 0;
        float bo[3], bd[3];
        for (uint32_t j = 0; j < ninst; ++j) {
            if (j == skip || !I[j].valid) continue;
            float o[3], d[3];
            for (int k = 0; k < 3; ++k) {
                if (cam) {
                    o[k] = I[j].eye_m[k];
                    d[k] = (I[j].dirm[0][k] * cam[0] + I[j].dirm[1][k] * cam[1]) + I[j].dirm[3][k];
                } else {
                    o[k] = ((I[j].Mi.c[0][k] * ow[0] + I[j].Mi.c[1][k] * ow[1]) + I[j].Mi.c[2][k] * ow[2]) + I[j].Mi.c[3][k];
                    d[k] = (I[j].Mi.c[0][k] * dw[0] + I[j].Mi.c[1][k] * dw[1]) + I[j].Mi.c[2][k] * dw[2];
                }
            }
            float tn;
            int axis;
            if (!slab_unit_cube(o, d, &tn, &axis)) continue;
            if (have_last && !(tn > last_t || (tn == last_t && j > last_j))) continue;
            if (!found || tn < best_t) { /* j ascending, so ties keep the lower index */
                found = 1; best_t = tn; best_j = j; best_axis = axis;
                memcpy(bo, o, sizeof bo); memcpy(bd, d, sizeof bd);
            }
        }
        if (!found) return;
        float mp[3], pos[3];
        entry_point(bo, bd, best_t, best_axis, mp);
        const inst_uniforms* J = &I[best_j];
        const float size[3] = {(float)(int32_t)J->w, (float)(int32_t)J->h, (float)(int32_t)J->d};
        for (int k = 0; k < 3; ++k) pos[k] = (mp[k] + 0.5f) * size[k];
        dda_march(&J->vol, pos, bd, NULL, &out->dda);
        *iters += out->dda.steps;
        if (out->dda.hit) {
            out->hit = 1; out->instance = best_j; out->entry_axis = best_axis;
            return;
        }
        last_t = best_t; last_j = best_j; have_last = 1;
    }
}

static void trace_path(const frame_uniforms* F, const inst_uniforms* I, uint32_t ninst, int px, int py,
                       uint32_t sample, uint32_t seed, uint32_t bounces, float L[3], uint64_t* rays,
                       uint64_t* iters) {
    vo_rng rng;
    rng_init(&rng, seed, (uint32_t)py * (uint32_t)F->width + (uint32_t)px, sample);
    float jx = rng_u01(&rng), jy = rng_u01(&rng);
    float fx = (float)px + jx, fy = (float)py + jy;
    /* camera segment: the ray from the eye through the jittered sample position, traced with the
     * same visibility rule as every later segment (nearest box entry first, alpha > 0 = opaque;
     * a path tracer has no near/far clip planes and no proxy-depth test) */
    path_hit cur;
    const float cam[2] = {fx * F->sxn - 1.0f, fy * F->syn - 1.0f};
    trace_world(I, ninst, 0xFFFFFFFFu, cam, NULL, NULL, &cur, iters);
    *rays += 1;
    const float sky[3] = {53.0f / 100.0f, 81.0f / 100.0f, 92.0f / 100.0f}; /* lib/command.c:57-59 */
    float thr[3] = {1.0f, 1.0f, 1.0f};
    L[0] = L[1] = L[2] = 0.0f;
    for (uint32_t b = 0;; ++b) {
        if (!cur.hit) {
            for (int c = 0; c < 3; ++c) L[c] = thr[c] * sky[c];
            return;
        }
        for (int c = 0; c < 3; ++c) thr[c] = thr[c] * g_srgb_dec[cur.dda.rgba[c]];
        if (b == bounces) return;
        const inst_uniforms* J = &I[cur.instance];
        const dda_state* r = &cur.dda;
        const float size[3] = {(float)(int32_t)J->w, (float)(int32_t)J->h, (float)(int32_t)J->d};
        int a, nsign;
        float p0[3];
        int32_t sv[3];
        leave_hit(r, cur.entry_axis, size, &a, &nsign, p0, sv);
        /* cosine-weighted direction about the face normal: normalize(n + uniform sphere point) */
        float dn[3];
        rng_sphere(&rng, dn);
        dn[a] += (float)nsign;
        float l2 = (dn[0] * dn[0] + dn[1] * dn[1]) + dn[2] * dn[2];
        if (l2 < 1e-6f) {
            dn[0] = dn[1] = dn[2] = 0.0f;
            dn[a] = (float)nsign;
        } else {
            float rl = 1.0f / sqrtf(l2);
            dn[0] *= rl; dn[1] *= rl; dn[2] *= rl;
        }
        *rays += 1;
        path_hit next;
        next.hit = 0;
        next.instance = cur.instance;
        next.entry_axis = a;
        dda_march(&J->vol, p0, dn, sv, &next.dda);
        *iters += next.dda.steps;
        next.hit = next.dda.hit;
        if (!next.hit && ninst > 1) {
            float pm[3], dm[3], ow[3], dw[3];
            for (int k = 0; k < 3; ++k) { pm[k] = p0[k] / size[k] - 0.5f; dm[k] = dn[k] / size[k]; }
            for (int k = 0; k < 3; ++k) {
                ow[k] = ((J->M.c[0][k] * pm[0] + J->M.c[1][k] * pm[1]) + J->M.c[2][k] * pm[2]) + J->M.c[3][k];
                dw[k] = (J->M.c[0][k] * dm[0] + J->M.c[1][k] * dm[1]) + J->M.c[2][k] * dm[2];
            }
            trace_world(I, ninst, cur.instance, NULL, ow, dw, &next, iters);
        }
        cur = next;
    }
}

void vo_render_paths(const vo_scene* s, const float* P, const float* V, int width, int height,
                     uint32_t flags, uint32_t bounces, uint32_t seed, uint32_t sample_first,
                     uint32_t sample_stride, uint32_t sample_count, uint64_t* accum, uint64_t* stats,
                     int num_threads) {
    srgb_init();
    frame_uniforms F;
    frame_setup(&F, P, V, width, height, flags);
    inst_uniforms* I = (inst_uniforms*)malloc(sizeof(inst_uniforms) * s->ninst);
    for (uint32_t i = 0; i < s->ninst; ++i) inst_setup(&I[i], &F, s, s->inst + 16 * (size_t)i);
    uint64_t rays = 0, iters = 0;
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads) reduction(+ : rays, iters)
#endif
    for (int py = 0; py < height; ++py) {
        for (int px = 0; px < width; ++px) {
            uint64_t acc[3] = {0, 0, 0};
            for (uint32_t k = 0; k < sample_count; ++k) {
                float L[3];
                uint64_t r = 0, it = 0;
                trace_path(&F, I, s->ninst, px, py, sample_first + k * sample_stride, seed, bounces, L, &r, &it);
                rays += r; iters += it;
                for (int c = 0; c < 3; ++c) {
                    float q = L[c] * 16777216.0f;
                    acc[c] += (q == q && q > 0.0f) ? (uint64_t)q : 0u;
                }
            }
            size_t p = (size_t)py * (size_t)width + (size_t)px;
            for (int c = 0; c < 3; ++c) accum[3 * p + c] += acc[c];
        }
    }
    free(I);
    if (stats) { stats[0] += rays; stats[1] += iters; }
}

/* Incoherent-ray extension: n rays through the volume of instance 0, in its
 * voxel space; ray i has origin uniform in the volume and direction uniform on the sphere, both from
 * the RNG stream keyed (seed, i, 0).  Record: hit_voxel = x | y << 16, instance = z, packed as for
 * primary rays (entry axis bit = 0 when steps == 0), iters = steps; VO_MISS / 0 / VO_MISS on a miss. */
uint64_t vo_render_rays(const vo_scene* s, uint64_t n, uint64_t first, uint32_t seed, vo_hit_record* records, uint8_t* rgba8,
                        int num_threads) {
    srgb_init();
    uint32_t tex;
    memcpy(&tex, s->inst + 15, 4);
    if (tex >= s->ntex) return 0;
    const vo_texture* t = &s->tex[tex];
    const vol_view vol = {t->kind, t->w, t->h, t->d, t->seed, t->rgba, t->heights, t->btable, t->bmasks, t->bcolors};
    const float size[3] = {(float)(int32_t)t->w, (float)(int32_t)t->h, (float)(int32_t)t->d};
    uint64_t total = 0;
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(num_threads) reduction(+ : total)
#endif
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        const uint64_t i = first + (uint64_t)k;
        vo_rng rng;
        rng_init(&rng, seed, (uint32_t)i, (uint32_t)(i >> 32));
        float pos[3], dir[3];
        for (int c = 0; c < 3; ++c) pos[c] = rng_u01(&rng) * size[c];
        rng_sphere(&rng, dir);
        dda_state r;
        dda_march(&vol, pos, dir, NULL, &r);
        total += r.steps;
        vo_hit_record rec = {VO_MISS, 0, VO_MISS, r.steps};
        uint8_t px[4] = {0, 0, 0, 0};
        if (r.hit) {
            rec.hit_voxel = (uint32_t)r.voxel[0] | ((uint32_t)r.voxel[1] << 16);
            rec.instance = (uint32_t)r.voxel[2];
            const uint32_t neg = (uint32_t)(r.step[0] < 0) | ((uint32_t)(r.step[1] < 0) << 1) | ((uint32_t)(r.step[2] < 0) << 2);
            rec.packed = (r.steps & 0xFFFFu) | ((r.last_mask | (neg << 3)) << 16);
            memcpy(px, r.rgba, 4);
        }
        if (records) records[k] = rec;
        if (rgba8) memcpy(rgba8 + 4 * (size_t)k, px, 4);
    }
    return total;
}

void vo_resolve(const uint64_t* accum, int width, int height, uint32_t total_spp, uint8_t* rgba8) {
    srgb_init();
    float scale = 1.0f / ((float)total_spp * 16777216.0f);
    size_t n = (size_t)width * (size_t)height;
    for (size_t p = 0; p < n; ++p) {
        for (int c = 0; c < 3; ++c) rgba8[4 * p + c] = vo_srgb_encode((float)accum[3 * p + c] * scale);
        rgba8[4 * p + 3] = 255;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* .vox loader restatement: src/voxel/magica_voxel.rs:18-44 over dot_vox 4.1.0                */
/* (Cargo.lock:127-130; un-vendored).  dot_vox semantics restated from its published format   */
/* handling: RIFF-like chunks under MAIN; SIZE then XYZI per model; voxel.i = file index - 1; */
/* palette[k] = k-th RGBA quad of the RGBA chunk as a little-endian u32 (default palette when */
/* the chunk is absent — not needed for the two assets, which both carry RGBA).  IMAP, nTRN,  */
/* nGRP, nSHP, MATL, LAYR, rOBJ ... are skipped.                                              */

static uint32_t rd_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* MagicaVoxel's default palette (the table published with the .vox format: index 0 unused, a 6x6x6 colour cube without
 * black, then ramps of red, green, blue and grey), laid out like an RGBA chunk: entry k is the colour of file index k + 1.
 * dot_vox 4.1.0 supplies its copy of this table when a file has no RGBA chunk; that copy (resources/default_palette.bytes
 * of the crate) is not available offline, so this restatement is unpinned. */
static void vox_default_palette(uint8_t pal[1024]) {
    static const uint8_t lv[6] = {0xff, 0xcc, 0x99, 0x66, 0x33, 0x00};
    static const uint8_t ramp[10] = {0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11};
    uint32_t t[257];
    t[0] = 0;
    for (uint32_t k = 0; k < 215; ++k) /* 0xAABBGGRR: blue runs fastest, red slowest */
        t[k + 1] = 0xff000000u | ((uint32_t)lv[k % 6] << 16) | ((uint32_t)lv[(k / 6) % 6] << 8) | (uint32_t)lv[k / 36];
    for (uint32_t j = 0; j < 10; ++j) {
        t[216 + j] = 0xff000000u | ramp[j];
        t[226 + j] = 0xff000000u | ((uint32_t)ramp[j] << 8);
        t[236 + j] = 0xff000000u | ((uint32_t)ramp[j] << 16);
        t[246 + j] = 0xff000000u | ((uint32_t)ramp[j] * 0x010101u);
    }
    t[256] = 0;
    for (uint32_t k = 0; k < 256; ++k) {
        const uint32_t c = t[k + 1];
        pal[4 * k + 0] = (uint8_t)c; pal[4 * k + 1] = (uint8_t)(c >> 8); pal[4 * k + 2] = (uint8_t)(c >> 16); pal[4 * k + 3] = (uint8_t)(c >> 24);
    }
}

/* model `model` of the file (dot_vox: one model per SIZE / XYZI pair, in file order); -6 when the file has fewer */
int vo_load_vox_model(const char* path, uint32_t model, uint32_t dims[3], uint8_t* out, uint64_t out_capacity) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc((size_t)n);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return -2; }
    fclose(f);
    int rc = -3;
    if (n < 20 || memcmp(buf, "VOX ", 4) != 0 || memcmp(buf + 8, "MAIN", 4) != 0) goto done;
    {
        const uint8_t* size_chunk = NULL;
        const uint8_t* xyzi_chunk = NULL;
        const uint8_t* rgba_chunk = NULL;
        const uint8_t* pending_size = NULL;
        uint32_t seen = 0;
        uint8_t default_pal[1024];
        size_t off = 20; /* "VOX " ver "MAIN" n m */
        while (off + 12 <= (size_t)n) {
            const uint8_t* id = buf + off;
            uint32_t cn = rd_u32(buf + off + 4), cm = rd_u32(buf + off + 8);
            const uint8_t* body = buf + off + 12;
            if (off + 12 + (size_t)cn + (size_t)cm > (size_t)n) break;
            if (!memcmp(id, "SIZE", 4)) pending_size = body;
            else if (!memcmp(id, "XYZI", 4) && pending_size) {
                if (seen == model) { size_chunk = pending_size; xyzi_chunk = body; }
                ++seen;
                pending_size = NULL;
            } else if (!memcmp(id, "RGBA", 4) && !rgba_chunk) rgba_chunk = body;
            off += 12 + (size_t)cn + (size_t)cm;
        }
        if (!size_chunk || !xyzi_chunk) { rc = seen ? -6 : -3; goto done; }
        if (!rgba_chunk) { vox_default_palette(default_pal); rgba_chunk = default_pal; }
        uint32_t sx = rd_u32(size_chunk), sy = rd_u32(size_chunk + 4), sz = rd_u32(size_chunk + 8);
        dims[0] = sx; dims[1] = sy; dims[2] = sz; /* RawDynamicChunk::new(size.x, size.y, size.z), :23-28 */
        rc = 0;
        if (!out) goto done;
        uint64_t bytes = (uint64_t)4 * sx * sy * sz;
        if (out_capacity < bytes) { rc = -4; goto done; }
        memset(out, 0, (size_t)bytes); /* Color::from_uint(0), :27 */
        uint32_t nv = rd_u32(xyzi_chunk);
        for (uint32_t v = 0; v < nv; ++v) {
            const uint8_t* q = xyzi_chunk + 4 + 4 * (size_t)v;
            int32_t vx = q[0], vy = q[1], vz = q[2];
            uint32_t i = q[3] ? (uint32_t)q[3] - 1u : 0u; /* dot_vox: i = index - 1 */
            /* chunk.at_mut(voxel.x, size.y - voxel.z - 1, voxel.y), :31-37 */
            int32_t cx = vx, cy = (int32_t)sy - vz - 1, cz = vy;
            if (cx < 0 || cy < 0 || cz < 0 || cx >= (int32_t)sx || cy >= (int32_t)sy || cz >= (int32_t)sz) {
                rc = -5; /* .unwrap() on None would panic in the reference */
                goto done;
            }
            /* data[z + dim_z*(y + dim_y*x)], src/voxel/rawchunk.rs:292 ; Color{r,g,b,a} = LE bytes */
            size_t idx = (size_t)cz + (size_t)sz * ((size_t)cy + (size_t)sy * (size_t)cx);
            memcpy(out + 4 * idx, rgba_chunk + 4 * (size_t)i, 4);
        }
    }
done:
    free(buf);
    return rc;
}

int vo_load_vox(const char* path, uint32_t dims[3], uint8_t* out, uint64_t out_capacity) {
    return vo_load_vox_model(path, 0, dims, out, out_capacity);
}

/* texels of volume `tex` (any kind) in the box [x0,x0+nx) x [y0,y0+ny) x [z0,z0+nz), x fastest: lets the tests turn a
 * procedural or brick volume into the equivalent dense texture and check that both traverse identically */
int vo_read_texels(const vo_scene* s, uint32_t tex, uint32_t x0, uint32_t y0, uint32_t z0, uint32_t nx, uint32_t ny, uint32_t nz,
                   uint8_t* rgba) {
    if (tex >= s->ntex) return -1;
    const vo_texture* t = &s->tex[tex];
    if (x0 + nx > t->w || y0 + ny > t->h || z0 + nz > t->d) return -1;
    const vol_view vol = {t->kind, t->w, t->h, t->d, t->seed, t->rgba, t->heights, t->btable, t->bmasks, t->bcolors};
    for (uint32_t z = 0; z < nz; ++z)
        for (uint32_t y = 0; y < ny; ++y)
            for (uint32_t x = 0; x < nx; ++x)
                vol_texel(&vol, (int32_t)(x0 + x), (int32_t)(y0 + y), (int32_t)(z0 + z), rgba + 4 * ((size_t)(z * ny + y) * nx + x));
    return 0;
}

int vo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
