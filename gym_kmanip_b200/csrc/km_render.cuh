// km_render.cuh -- camera observations of the Vision ids: a batched ray caster for the primitives of the completed model.
//
// Replaces, for the batched path, reference gym_kmanip/env_sim.py:140-145 (get_observation: physics.render(height,
// width, camera_id) per camera) and env_sim.py:187-188 (k_render).  The reference renders MuJoCo's scene with OpenGL; the
// scene's visual meshes are absent from the snapshot, so what is drawn is the completed model
// (assets/completion_spec.json): the table plane, the cube box, the finger-pad spheres and one capsule per moving link
// (parent link origin -> link origin; a link-mounted camera does not see the proxies that end at its own link).  Shading follows MuJoCo's fixed-function conventions (Blinn-Phong per light:
// headlight at the camera + the directional lights of scene.xml:10-12, material = geom rgba, specular 0.5, shininess
// 0.5 * 128; highlight terms below 1e-6 are dropped), evaluated per pixel; no shadows, fog or reflections.  Camera frames are MuJoCo's: the camera looks along
// -z with +y up, `mode="targetbody"` turns z away from the target body and x orthogonal to z and world up
// (mj_camlight), vertical field of view `fovy`, pixel centres at half-integers, row 0 on top.
//
// Two kernels:
//   k_render_setup<S,T,G>  per env (lane group): state -> forward kinematics -> one record of floats
//                          [camera origin + axes | primitive list]  (KM_REC_HDR + 16 floats per primitive)
//   k_render_pixels        scene-independent, one CTA per 32 x 32 pixel tile per env: record -> shared memory, warp 0
//                          culls the primitives against the tile's cone, 128 threads shade one pixel column of eight rows
//                          each into a shared-memory tile (tiles that see no primitive: table plane / background loop
//                          only), the tile goes out as 16-byte streaming stores (rows of 96 bytes).
// The pixel kernel's algorithmic HBM traffic is the image itself (W*H*3 bytes per env) plus one ~1 KB record.
#pragma once
#include "km_model.cuh"

namespace km {

enum { KM_REC_HDR = 16, KM_PRIM_FLOATS = 16, KM_RENDER_MAXPRIM = 32, KM_RENDER_TILE = 32 };
enum { KM_PRIM_SPHERE = 0, KM_PRIM_CAPSULE = 1, KM_PRIM_BOX = 2 };
enum { KM_MAT_TABLE = 0, KM_MAT_CUBE = 1, KM_MAT_LINK = 2, KM_MAT_PAD = 3 };

struct KmRenderParams {
  int W, H, tiles_x, tiles_y, rec_floats, nlight;
  float focal, tab_z, link_radius;
  float ambient[3], head_diffuse[3], head_specular[3];
  float ldir[4][3], ldiffuse[4][3], lspecular[4][3];   // ldir: unit vector TOWARDS the light
  float mat[4][3], mat_specular;
  int shin_squarings;                                  // shininess exponent 2^k
  float spec_cut;                                      // cosines below 2^(-20 / exponent) have no highlight: x^exponent < 1e-6 is dropped
  // camera: position in its link's frame (link < 0: world), tracked point likewise
  int cam_link, tgt_link;
  float cam_pos[3], tgt_pos[3];
};

template <class S> constexpr int render_nprim() {
  int n = 1 + S::NPAD;
  for (int l = 0; l < S::NVA; l++) n += S::dof_parent[l] >= 0 ? 1 : 0;
  return n;
}
template <class S> constexpr int render_rec_floats() { return KM_REC_HDR + KM_PRIM_FLOATS * render_nprim<S>(); }

#if defined(__CUDACC__)

// record of one env from the link frames of its stored state; e must hold kinematics() output
template <class S, typename T, int G, class E> __device__ void render_record(const E& e, const Model<S, T>& m, const Grp<G>& g,
                                                                              const KmRenderParams& P, float* rec) {
  typedef Dim<S> D;
  static_assert(render_nprim<S>() <= KM_RENDER_MAXPRIM, "primitive list too long");
  if (g.lane == 0) {
    T o[3], t[3];
    if (P.cam_link >= 0) {
      const T cp[3] = {(T)P.cam_pos[0], (T)P.cam_pos[1], (T)P.cam_pos[2]};
      mulv3(o, e.xmat[P.cam_link], cp);
      for (int i = 0; i < 3; i++) o[i] += e.xpos[P.cam_link][i];
    } else for (int i = 0; i < 3; i++) o[i] = (T)P.cam_pos[i];
    if (P.tgt_link >= 0) {
      const T tp[3] = {(T)P.tgt_pos[0], (T)P.tgt_pos[1], (T)P.tgt_pos[2]};
      mulv3(t, e.xmat[P.tgt_link], tp);
      for (int i = 0; i < 3; i++) t[i] += e.xpos[P.tgt_link][i];
    } else for (int i = 0; i < 3; i++) t[i] = (T)P.tgt_pos[i];
    // mj_camlight, targetbody: z = normalize(cam - target), x = normalize(up x z), y = normalize(z x x)
    T z[3] = {o[0] - t[0], o[1] - t[1], o[2] - t[2]}, up[3] = {0, 0, 1}, x[3], y[3];
    normalize3(z);
    cross3(x, up, z);
    normalize3(x);
    cross3(y, z, x);
    normalize3(y);
    for (int i = 0; i < 3; i++) { rec[i] = (float)o[i]; rec[3 + i] = (float)x[i]; rec[6 + i] = (float)y[i]; rec[9 + i] = (float)z[i]; }
    constexpr int NP = render_nprim<S>();
    rec[12] = __int_as_float(NP);
    rec[13] = 0; rec[14] = 0; rec[15] = 0;
    // cube
    float* p = rec + KM_REC_HDR;
    p[0] = __int_as_float(KM_PRIM_BOX | (KM_MAT_CUBE << 4));
    for (int i = 0; i < 3; i++) { p[1 + i] = (float)(e.qpos[D::NVA + i] + e.cube_lo[i]); p[4 + i] = (float)m.cube_size[i]; }
    for (int i = 0; i < 9; i++) p[7 + i] = (float)e.cmat[i];
  }
  KM_FOR(s, D::NPAD) {
    const int l = m.pad_link[s];
    T c[3];
    mulv3(c, e.xmat[l], m.pad_pos[s]);
    float* p = rec + KM_REC_HDR + KM_PRIM_FLOATS * (1 + s);
    p[0] = __int_as_float(KM_PRIM_SPHERE | (KM_MAT_PAD << 4));
    for (int i = 0; i < 3; i++) p[1 + i] = (float)(c[i] + e.xpos[l][i]);
    p[4] = (float)m.pad_rad[s];
    for (int i = 5; i < 16; i++) p[i] = 0;
  }
  KM_FOR(l, D::NVA) {
    const int par = m.parent[l];
    if (par < 0) continue;
    int idx = 0;
    for (int k = 0; k < l; k++) idx += m.parent[k] >= 0 ? 1 : 0;
    float* p = rec + KM_REC_HDR + KM_PRIM_FLOATS * (1 + D::NPAD + idx);
    const T* a = e.xpos[par];
    const T* b = e.xpos[l];
    const T d2 = (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
    // coincident origins (e.g. a wrist pair): the capsule degenerates to a sphere.  A camera does not see the proxies
    // that end at its own link (they would wall it in): those become spheres of radius 0, which no ray hits.
    const bool hidden = P.cam_link >= 0 && (l == P.cam_link || par == P.cam_link);
    p[0] = __int_as_float((hidden || d2 < T(1e-10) ? KM_PRIM_SPHERE : KM_PRIM_CAPSULE) | (KM_MAT_LINK << 4));
    for (int i = 0; i < 3; i++) { p[1 + i] = (float)a[i]; p[5 + i] = (float)b[i]; }
    p[4] = hidden ? 0.0f : P.link_radius;
    for (int i = 8; i < 16; i++) p[i] = 0;
  }
}

#ifdef KM_RENDER_PIXELS_IMPL

__device__ __forceinline__ float rdot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void rnorm(float* a) {
  const float s = rsqrtf(fmaxf(rdot(a, a), 1e-30f));
  a[0] *= s; a[1] *= s; a[2] *= s;
}
// unit ray direction through pixel-space point (u, v) (pixel centres at half-integers)
__device__ __forceinline__ void pixel_ray(const KmRenderParams& P, const float* rec, float u, float v, float* d) {
  const float inv = 1.0f / P.focal;
  const float dx = (u - 0.5f * (float)P.W) * inv, dy = -(v - 0.5f * (float)P.H) * inv;
  for (int i = 0; i < 3; i++) d[i] = rec[3 + i] * dx + rec[6 + i] * dy - rec[9 + i];
  rnorm(d);
}

// nearest hit of the ray (o, d) with primitive p closer than tbest; on a hit updates tbest, the normal and the material
__device__ __forceinline__ void hit_prim(const float* p, const float* o, const float* d, float& tbest, float* nrm, int& mat) {
  const int tm = __float_as_int(p[0]), type = tm & 15;
  if (type == KM_PRIM_BOX) {
    const float* R = p + 7;
    const float oc[3] = {o[0] - p[1], o[1] - p[2], o[2] - p[3]};
    float ol[3], dl[3];
    for (int i = 0; i < 3; i++) { ol[i] = R[i] * oc[0] + R[3 + i] * oc[1] + R[6 + i] * oc[2]; dl[i] = R[i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2]; }
    float tn = -3.0e38f, tf = 3.0e38f;
    int ax = 0;
    for (int i = 0; i < 3; i++) {
      const float di = fabsf(dl[i]) < 1e-12f ? (dl[i] < 0 ? -1e-12f : 1e-12f) : dl[i];
      const float inv = 1.0f / di, t1 = (-ol[i]) * inv - fabsf(inv) * p[4 + i], t2 = (-ol[i]) * inv + fabsf(inv) * p[4 + i];
      if (t1 > tn) { tn = t1; ax = i; }
      tf = fminf(tf, t2);
    }
    if (tn <= tf && tn > 0.0f && tn < tbest) {
      tbest = tn;
      const float sg = dl[ax] > 0 ? -1.0f : 1.0f;
      for (int i = 0; i < 3; i++) nrm[i] = sg * R[3 * i + ax];
      mat = tm >> 4;
    }
    return;
  }
  const float r = p[4];
  const float oa[3] = {o[0] - p[1], o[1] - p[2], o[2] - p[3]};
  if (type == KM_PRIM_SPHERE) {
    const float b = rdot(d, oa), c = rdot(oa, oa) - r * r, h = b * b - c;
    if (h > 0.0f) {
      const float t = -b - sqrtf(h);
      if (t > 0.0f && t < tbest) {
        tbest = t;
        const float ir = 1.0f / r;
        for (int i = 0; i < 3; i++) nrm[i] = (oa[i] + t * d[i]) * ir;
        mat = tm >> 4;
      }
    }
    return;
  }
  // capsule a = p[1..3], b = p[5..7]
  const float ba[3] = {p[5] - p[1], p[6] - p[2], p[7] - p[3]};
  const float baba = rdot(ba, ba), bard = rdot(ba, d), baoa = rdot(ba, oa), rdoa = rdot(d, oa), oaoa = rdot(oa, oa);
  const float a = baba - bard * bard, b = baba * rdoa - baoa * bard, c = baba * oaoa - baoa * baoa - r * r * baba;
  const float h = b * b - a * c;
  if (h < 0.0f) return;
  float t = (-b - sqrtf(h)) / a;
  const float y = baoa + t * bard;
  float hh = y / baba;
  if (!(y > 0.0f && y < baba)) {
    const float oc[3] = {y <= 0.0f ? oa[0] : o[0] - p[5], y <= 0.0f ? oa[1] : o[1] - p[6], y <= 0.0f ? oa[2] : o[2] - p[7]};
    const float b2 = rdot(d, oc), c2 = rdot(oc, oc) - r * r, h2 = b2 * b2 - c2;
    if (!(h2 > 0.0f)) return;
    t = -b2 - sqrtf(h2);
    hh = y <= 0.0f ? 0.0f : 1.0f;
  }
  if (t > 0.0f && t < tbest) {
    tbest = t;
    const float ir = 1.0f / r;
    for (int i = 0; i < 3; i++) nrm[i] = (oa[i] + t * d[i] - hh * ba[i]) * ir;
    mat = tm >> 4;
  }
}

// x^(2^k) by repeated squaring.  KS >= 0: k is the compile-time constant KS (MuJoCo's default shininess 0.5 -> exponent 64,
// k = 6: the instantiation every shipped scene uses); KS < 0: run-time k
template <int KS> __device__ __forceinline__ float shin_pow(float x, int k) {
  if constexpr (KS >= 0) {
#pragma unroll
    for (int i = 0; i < KS; i++) x *= x;
    return x;
  } else {
#pragma unroll 1
    for (int i = 0; i < k; i++) x *= x;
    return x;
  }
}

// Blinn-Phong of one surface point: ambient + headlight (at the camera) + directional lights; V = unit vector to the camera
template <int KS> __device__ __forceinline__ void shade(const KmRenderParams& P, const float* N, const float* V, int mat, float* rgb) {
  float dif[3] = {P.ambient[0], P.ambient[1], P.ambient[2]}, spc[3] = {0, 0, 0};
  const float nv = fmaxf(rdot(N, V), 0.0f);
  for (int c = 0; c < 3; c++) dif[c] += P.head_diffuse[c] * nv;
  if (nv > P.spec_cut) {   // warp-coherent in practice: highlights are compact blobs
    const float s = shin_pow<KS>(nv, P.shin_squarings);
    for (int c = 0; c < 3; c++) spc[c] += P.head_specular[c] * s;
  }
#pragma unroll
  for (int l = 0; l < 4; l++) {   // fixed trip count: the light constants become immediate constant-bank operands
    if (l < P.nlight) {
      const float nl = rdot(N, P.ldir[l]);
      if (nl > 0.0f) {
        float Hh[3] = {P.ldir[l][0] + V[0], P.ldir[l][1] + V[1], P.ldir[l][2] + V[2]};
        rnorm(Hh);
        const float nh = rdot(N, Hh);
        for (int c = 0; c < 3; c++) dif[c] += P.ldiffuse[l][c] * nl;
        if (nh > P.spec_cut) {
          const float sh = shin_pow<KS>(nh, P.shin_squarings);
          for (int c = 0; c < 3; c++) spc[c] += P.lspecular[l][c] * sh;
        }
      }
    }
  }
  for (int c = 0; c < 3; c++) rgb[c] = fminf(P.mat[mat][c] * dif[c] + P.mat_specular * spc[c], 1.0f);
}

// The same for a point of the table plane (most pixels of every camera): the normal is +z, so every N.L is a constant of
// the launch (`tdif` = ambient + sum of the lights' diffuse terms, formed once per thread; `lmask` = the lights above the
// plane) and the half vector needs no normalisation of its own: |L + V|^2 = 2 + 2 L.V for unit L and V, hence
// N.H = (L_z + V_z) rsqrt(2 + 2 L.V).  About a third of the instructions of shade(); rounding differs from it by an ulp
// here and there (one grey level at most).
template <int KS> __device__ __forceinline__ void shade_table(const KmRenderParams& P, const float* d, const float* tdif, unsigned lmask, float* rgb) {
  const float vz = -d[2], nv = fmaxf(vz, 0.0f);
  float dif[3] = {tdif[0], tdif[1], tdif[2]}, spc[3];
  for (int c = 0; c < 3; c++) dif[c] += P.head_diffuse[c] * nv;
  // highlights without branches (on the table they are on for most pixels): a select instead of divergence bookkeeping
  const float s = nv > P.spec_cut ? shin_pow<KS>(nv, P.shin_squarings) : 0.0f;
  for (int c = 0; c < 3; c++) spc[c] = P.head_specular[c] * s;
#pragma unroll
  for (int l = 0; l < 4; l++) {
    if ((lmask >> l) & 1u) {   // warp-uniform
      const float lv = rdot(P.ldir[l], d);
      const float nh = (P.ldir[l][2] + vz) * rsqrtf(fmaxf(2.0f - 2.0f * lv, 1e-12f));
      const float sh = nh > P.spec_cut ? shin_pow<KS>(nh, P.shin_squarings) : 0.0f;
      for (int c = 0; c < 3; c++) spc[c] += P.lspecular[l][c] * sh;
    }
  }
  for (int c = 0; c < 3; c++) rgb[c] = fminf(P.mat[KM_MAT_TABLE][c] * dif[c] + P.mat_specular * spc[c], 1.0f);
}

// One CTA of 128 threads per 32 x 32 pixel tile (blockIdx.x, blockIdx.y) of env env0 + blockIdx.z; every thread owns one
// pixel column of the tile and walks eight rows, so the horizontal part of the ray and everything that depends on the
// launch or the env only (camera axes, light terms of the table plane) is formed once per thread.  Tiles that see no
// primitive (most of them) take a loop that knows only the table plane and the background.
constexpr int KM_RENDER_THREADS = 128, KM_RENDER_ROWS = KM_RENDER_TILE * KM_RENDER_TILE / KM_RENDER_THREADS;
template <int KS> __global__ void __launch_bounds__(KM_RENDER_THREADS) k_render_pixels(const float* __restrict__ recs, unsigned char* __restrict__ out, KmRenderParams P, int env0) {
  __shared__ float rec[KM_REC_HDR + KM_PRIM_FLOATS * KM_RENDER_MAXPRIM];
  __shared__ unsigned s_mask;
  __shared__ __align__(16) unsigned char tile[KM_RENDER_TILE * KM_RENDER_TILE * 3];
  const int tid = threadIdx.x, env = env0 + blockIdx.z;
  const int tx0 = blockIdx.x * KM_RENDER_TILE, ty0 = blockIdx.y * KM_RENDER_TILE;
  const float* src = recs + (size_t)env * P.rec_floats;
  for (int i = tid; i < P.rec_floats; i += KM_RENDER_THREADS) rec[i] = src[i];
  __syncthreads();
  if (tid < 32) {
    // cull: keep the primitives whose bounding sphere meets the cone around the tile (axis = ray through the tile
    // centre, half-angle = the widest corner ray)
    const int nprim = __float_as_int(rec[12]);
    float axis[3], cr[3];
    pixel_ray(P, rec, (float)tx0 + 16.0f, (float)ty0 + 16.0f, axis);
    float cos_t = 1.0f;
    for (int k = 0; k < 4; k++) {
      pixel_ray(P, rec, (float)(tx0 + (k & 1) * KM_RENDER_TILE), (float)(ty0 + (k >> 1) * KM_RENDER_TILE), cr);
      cos_t = fminf(cos_t, rdot(axis, cr));
    }
    bool on = false;
    if (tid < nprim) {
      const float* p = rec + KM_REC_HDR + KM_PRIM_FLOATS * tid;
      const int type = __float_as_int(p[0]) & 15;
      float c[3] = {p[1], p[2], p[3]}, r;
      if (type == KM_PRIM_BOX) r = sqrtf(p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);
      else if (type == KM_PRIM_SPHERE) r = p[4];
      else {
        const float ba[3] = {p[5] - p[1], p[6] - p[2], p[7] - p[3]};
        for (int i = 0; i < 3; i++) c[i] += 0.5f * ba[i];
        r = 0.5f * sqrtf(rdot(ba, ba)) + p[4];
      }
      r *= 1.001f;
      const float v[3] = {c[0] - rec[0], c[1] - rec[1], c[2] - rec[2]};
      const float d2 = rdot(v, v);
      if (d2 <= r * r) on = true;
      else {
        const float id = rsqrtf(d2), ca = rdot(axis, v) * id, sin_s = r * id;
        const float cos_s = sqrtf(fmaxf(1.0f - sin_s * sin_s, 0.0f)), sin_t = sqrtf(fmaxf(1.0f - cos_t * cos_t, 0.0f));
        // inside when angle(axis, v) <= theta_t + theta_s; if the sum passes a right angle keep the primitive
        const float cs = cos_t * cos_s - sin_t * sin_s;
        on = cs <= 0.0f || ca >= cs - 1e-4f;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, on);
    if (tid == 0) s_mask = mask;
  }
  __syncthreads();
  const unsigned mask = s_mask;
  const float o[3] = {rec[0], rec[1], rec[2]};
  // per-thread constants: light terms of the table plane (normal +z), this pixel column's part of the ray
  float tdif[3] = {P.ambient[0], P.ambient[1], P.ambient[2]};
  unsigned lmask = 0;
#pragma unroll
  for (int l = 0; l < 4; l++)
    if (l < P.nlight && P.ldir[l][2] > 0.0f) {
      lmask |= 1u << l;
      for (int c = 0; c < 3; c++) tdif[c] += P.ldiffuse[l][c] * P.ldir[l][2];
    }
  const int lx = tid & 31, px = tx0 + lx, ly0 = tid >> 5;
  const float inv = 1.0f / P.focal, dx = ((float)px + 0.5f - 0.5f * (float)P.W) * inv;
  const float ay[3] = {rec[6], rec[7], rec[8]};
  const float base[3] = {rec[3] * dx - rec[9], rec[4] * dx - rec[10], rec[5] * dx - rec[11]};
  const bool above = o[2] > P.tab_z;
  if (mask == 0u) {
    // table plane and background only
#pragma unroll 2
    for (int k = 0; k < KM_RENDER_ROWS; k++) {
      const int ly = ly0 + 4 * k, py = ty0 + ly;
      float rgb[3] = {0, 0, 0};
      if (px < P.W && py < P.H) {
        const float dy = -((float)py + 0.5f - 0.5f * (float)P.H) * inv;
        float d[3] = {base[0] + ay[0] * dy, base[1] + ay[1] * dy, base[2] + ay[2] * dy};
        rnorm(d);
        if (above && d[2] < 0.0f) shade_table<KS>(P, d, tdif, lmask, rgb);
      }
      unsigned char* t = tile + ly * (KM_RENDER_TILE * 3) + lx * 3;
      for (int c = 0; c < 3; c++) t[c] = (unsigned char)(int)(rgb[c] * 255.0f + 0.5f);
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < KM_RENDER_ROWS; k++) {
      const int ly = ly0 + 4 * k, py = ty0 + ly;
      float rgb[3] = {0, 0, 0};
      if (px < P.W && py < P.H) {
        const float dy = -((float)py + 0.5f - 0.5f * (float)P.H) * inv;
        float d[3] = {base[0] + ay[0] * dy, base[1] + ay[1] * dy, base[2] + ay[2] * dy}, nrm[3] = {0, 0, 1};
        rnorm(d);
        float tbest = 3.0e38f;
        int mat = -1;
        if (above && d[2] < 0.0f) { tbest = (P.tab_z - o[2]) / d[2]; mat = KM_MAT_TABLE; }
        for (unsigned mm = mask; mm; mm &= mm - 1) {
          const int pi = __ffs(mm) - 1;
          hit_prim(rec + KM_REC_HDR + KM_PRIM_FLOATS * pi, o, d, tbest, nrm, mat);
        }
        if (mat == KM_MAT_TABLE) shade_table<KS>(P, d, tdif, lmask, rgb);
        else if (mat >= 0) {
          const float V[3] = {-d[0], -d[1], -d[2]};
          shade<KS>(P, nrm, V, mat, rgb);
        }
      }
      unsigned char* t = tile + ly * (KM_RENDER_TILE * 3) + lx * 3;
      for (int c = 0; c < 3; c++) t[c] = (unsigned char)(int)(rgb[c] * 255.0f + 0.5f);
    }
  }
  __syncthreads();
  unsigned char* img = out + (size_t)env * P.W * P.H * 3;
  const int rows = min(KM_RENDER_TILE, P.H - ty0);
  if ((P.W * 3) % 16 == 0 && tx0 + KM_RENDER_TILE <= P.W) {
    // full-width tile on a 16-byte-aligned pitch: each row is 96 contiguous bytes = six 16-byte streaming stores
    for (int i = tid; i < rows * 6; i += KM_RENDER_THREADS) {
      const int row = i / 6, seg = i % 6;
      const uint4 v = *(const uint4*)(tile + row * 96 + seg * 16);
      __stcs((uint4*)(img + ((size_t)(ty0 + row) * P.W + tx0) * 3 + seg * 16), v);
    }
  } else {
    const int cols = min(KM_RENDER_TILE, P.W - tx0) * 3;
    for (int i = tid; i < rows * cols; i += KM_RENDER_THREADS) {
      const int row = i / cols, cb = i % cols;
      img[((size_t)(ty0 + row) * P.W + tx0) * 3 + cb] = tile[row * 96 + cb];
    }
  }
}

#endif  // KM_RENDER_PIXELS_IMPL
#endif  // __CUDACC__

}  // namespace km
