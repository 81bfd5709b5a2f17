/*
 * raymarch_ref.c -- plain-C restatement of the reference's CUDA ray-marching kernels.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows nr4seg/nerf/raymarching/src/raymarching.cu and pcg32.h of ethz-asl/ucsa_neural_rendering:
 *   pcg32_*                  pcg32.h:44-116            (seed / next_uint / next_float)
 *   ref_near_far             raymarching.cu:62-115
 *   ref_march_rays_train     raymarching.cu:138-307    (occupancy march, two passes)
 *   ref_composite_train_fwd  raymarching.cu:318-394
 *   ref_composite_train_bwd  raymarching.cu:408-487
 *   ref_march_rays           raymarching.cu:528-634    (inference wavefront step)
 *   ref_composite_rays       raymarching.cu:647-729
 *   ref_compact_rays         raymarching.cu:837-855
 *
 * Arithmetic notes that matter for bit-exactness against a GPU build of the reference:
 *   - nvcc contracts a*b+c into FMA; the expressions it fuses are written with fmaf() here;
 *   - the grid index expression uses a double literal (raymarching.cu:200-202): evaluated in double here too;
 *   - frexpf / exp2f / IEEE division behave identically on both sides;
 *   - __expf (composite kernels) is an approximation on the GPU: those outputs are compared to 1e-5, not bit-exactly.
 * The reference reserves output ranges with atomicAdd, so its ray order is arbitrary; this restatement (and the CUDA
 * rewrite) hand out offsets in ray order, which is one of the orders the reference can produce.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libraymarch_oracle.so oracle/raymarch_ref.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAX_STEPS 1024
#define SQRT3 1.73205080757f
#define DENSITY_THRESH 0.01f

static inline float min_stepsize(void) { return 2 * SQRT3 / MAX_STEPS; }
static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf(float x) { return copysignf(1.0f, x); }

/* ------------------------------------------------------------------ PCG32 */
typedef struct {
  uint64_t state, inc;
} pcg32_t;

static uint32_t pcg32_next_uint(pcg32_t* r) {
  const uint64_t old = r->state;
  r->state = old * 0x5851f42d4c957f2dULL + r->inc;
  const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  const uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
static void pcg32_seed(pcg32_t* r, uint64_t initstate, uint64_t initseq) {
  r->state = 0u;
  r->inc = (initseq << 1u) | 1u;
  pcg32_next_uint(r);
  r->state += initstate;
  pcg32_next_uint(r);
}
static float pcg32_next_float(pcg32_t* r) {
  union {
    uint32_t u;
    float f;
  } x;
  x.u = (pcg32_next_uint(r) >> 9) | 0x3f800000u;
  return x.f - 1.0f;
}
/* exported for the known-answer test */
void ref_pcg32_stream(uint64_t initstate, uint64_t initseq, int n, uint32_t* out_uint, float* out_float) {
  pcg32_t a, b;
  pcg32_seed(&a, initstate, initseq);
  pcg32_seed(&b, initstate, initseq);
  for (int i = 0; i < n; ++i) {
    out_uint[i] = pcg32_next_uint(&a);
    out_float[i] = pcg32_next_float(&b);
  }
}

/* ------------------------------------------------------------------ near / far */
void ref_near_far(const float* rays_o, const float* rays_d, const float* aabb, uint32_t n_rays, float min_near,
                  float* nears, float* fars) {
  for (uint32_t n = 0; n < n_rays; ++n) {
    const float* o = rays_o + 3 * n;
    const float* d = rays_d + 3 * n;
    float near = 0, far = 0;
    int hit = 1;
    for (int ax = 0; ax < 3 && hit; ++ax) {
      const float inv = 1 / d[ax];
      float lo = (aabb[ax] - o[ax]) * inv, hi = (aabb[3 + ax] - o[ax]) * inv;
      if (lo > hi) {
        const float t = lo;
        lo = hi;
        hi = t;
      }
      if (ax == 0) {
        near = lo;
        far = hi;
      } else if (near > hi || lo > far) {
        hit = 0;
      } else {
        if (lo > near) near = lo;
        if (hi < far) far = hi;
      }
    }
    if (!hit) {
      nears[n] = fars[n] = 3.402823466e+38f;
      continue;
    }
    if (near < min_near) near = min_near;
    nears[n] = near;
    fars[n] = far;
  }
}

/* ------------------------------------------------------------------ the marcher's inner step */
typedef struct {
  float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
  float bound, dt_gamma, dt_min, dt_max, thresh;
  uint32_t C, H;
  const float* grid;
} march_ctx;

static int mip_from_pos(float x, float y, float z, float max_cascade) {
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int e;
  frexpf(mx, &e);
  return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

/* one evaluation at parameter t: returns 1 when the cell is occupied (x,y,z valid), and advances *t either by one
 * step (occupied) or to the next voxel boundary (empty), exactly as raymarching.cu:187-226 */
static int march_probe(const march_ctx* c, float* t, float* x, float* y, float* z, float* dt_out) {
  *x = clampf(fmaf(*t, c->dx, c->ox), -c->bound, c->bound);
  *y = clampf(fmaf(*t, c->dy, c->oy), -c->bound, c->bound);
  *z = clampf(fmaf(*t, c->dz, c->oz), -c->bound, c->bound);
  const int level = mip_from_pos(*x, *y, *z, (float)c->C);
  const float mip_bound = fminf(exp2f((float)level), c->bound);
  const float mip_rbound = 1 / mip_bound;
  const float H = (float)c->H;
  /* 0.5 (double) * (x*rb + 1) (float, fused) * H (uint -> double); then double -> float for clamp, then int */
  const int nx = (int)clampf((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)c->H), 0.0f, H - 1);
  const int ny = (int)clampf((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)c->H), 0.0f, H - 1);
  const int nz = (int)clampf((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)c->H), 0.0f, H - 1);
  const uint32_t index = level * c->H * c->H * c->H + nx * c->H * c->H + ny * c->H + nz;
  const float density = c->grid[index];
  if (density > c->thresh) {
    const float dt = clampf(*t * c->dt_gamma, c->dt_min, c->dt_max);
    *t += dt;
    *dt_out = dt;
    return 1;
  }
  const float hm1 = (float)(c->H - 1);
  const float tx = (fmaf(fmaf(0.5f, signf(c->dx), nx + 0.5f) / hm1 * 2 - 1, mip_bound, -*x)) * c->rdx;
  const float ty = (fmaf(fmaf(0.5f, signf(c->dy), ny + 0.5f) / hm1 * 2 - 1, mip_bound, -*y)) * c->rdy;
  const float tz = (fmaf(fmaf(0.5f, signf(c->dz), nz + 0.5f) / hm1 * 2 - 1, mip_bound, -*z)) * c->rdz;
  const float tt = *t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
  do {
    const float dt = clampf(*t * c->dt_gamma, c->dt_min, c->dt_max);
    *t += dt;
  } while (*t < tt);
  return 0;
}

static void ctx_for_ray(march_ctx* c, const float* o, const float* d, const float* grid, float mean_density, float bound,
                        float dt_gamma, uint32_t C, uint32_t H) {
  c->ox = o[0], c->oy = o[1], c->oz = o[2];
  c->dx = d[0], c->dy = d[1], c->dz = d[2];
  c->rdx = 1 / d[0], c->rdy = 1 / d[1], c->rdz = 1 / d[2];
  c->bound = bound, c->dt_gamma = dt_gamma;
  c->dt_min = min_stepsize();
  c->dt_max = 2 * bound / H;
  c->thresh = fminf(DENSITY_THRESH, mean_density);
  c->C = C, c->H = H, c->grid = grid;
}

/* rays: [N,3] = (ray id, offset, count); counter[0] += total samples, counter[1] += N.  Offsets in ray order. */
void ref_march_rays_train(const float* rays_o, const float* rays_d, const float* grid, float mean_density, float bound,
                          float dt_gamma, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                          uint32_t perturb) {
  for (uint32_t n = 0; n < N; ++n) {
    march_ctx c;
    ctx_for_ray(&c, rays_o + 3 * n, rays_d + 3 * n, grid, mean_density, bound, dt_gamma, C, H);
    const float far = fars[n];
    float t0 = nears[n];
    if (perturb) {
      pcg32_t rng;
      pcg32_seed(&rng, (uint64_t)n, 1u);
      t0 = fmaf(min_stepsize(), pcg32_next_float(&rng), t0); /* nvcc fuses this (raymarching.cu:178) */
    }
    float t = t0, x, y, z, dt;
    uint32_t num_steps = 0;
    while (t < far && num_steps < MAX_STEPS) num_steps += march_probe(&c, &t, &x, &y, &z, &dt);
    const uint32_t point_index = (uint32_t)counter[0];
    const uint32_t ray_index = (uint32_t)counter[1];
    counter[0] += (int32_t)num_steps;
    counter[1] += 1;
    rays[ray_index * 3] = (int32_t)n;
    rays[ray_index * 3 + 1] = (int32_t)point_index;
    rays[ray_index * 3 + 2] = (int32_t)num_steps;
    if (num_steps == 0 || point_index + num_steps >= M) continue;
    float* px = xyzs + 3ull * point_index;
    float* pd = dirs + 3ull * point_index;
    float* pl = deltas + 2ull * point_index;
    t = t0;
    float last_t = t;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
      if (march_probe(&c, &t, &x, &y, &z, &dt)) {
        px[0] = x, px[1] = y, px[2] = z;
        pd[0] = c.dx, pd[1] = c.dy, pd[2] = c.dz;
        pl[0] = dt;
        pl[1] = t - last_t;
        last_t = t;
        px += 3, pd += 3, pl += 2;
        ++step;
      }
    }
  }
}

void ref_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t C, uint32_t H,
                    const float* grid, float mean_density, const float* nears, const float* fars, float* xyzs,
                    float* dirs, float* deltas, uint32_t perturb) {
  for (uint32_t n = 0; n < n_alive; ++n) {
    const int index = rays_alive[n];
    march_ctx c;
    ctx_for_ray(&c, rays_o + 3 * index, rays_d + 3 * index, grid, mean_density, bound, dt_gamma, C, H);
    const float far = fars[index];
    float t = rays_t[n];
    if (perturb) {
      pcg32_t rng;
      pcg32_seed(&rng, (uint64_t)n, (uint64_t)perturb);
      t = fmaf(min_stepsize(), pcg32_next_float(&rng), t); /* raymarching.cu:574, fused */
    }
    float* px = xyzs + 3ull * n * n_step;
    float* pd = dirs + 3ull * n * n_step;
    float* pl = deltas + 2ull * n * n_step;
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
      if (march_probe(&c, &t, &x, &y, &z, &dt)) {
        px[0] = x, px[1] = y, px[2] = z;
        pd[0] = c.dx, pd[1] = c.dy, pd[2] = c.dz;
        pl[0] = dt;
        pl[1] = t - last_t;
        last_t = t;
        px += 3, pd += 3, pl += 2;
        ++step;
      }
    }
  }
}

/* ------------------------------------------------------------------ ragged compositing (training) */
void ref_composite_train_fwd(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                             uint32_t M, uint32_t N, float* weights_sum, float* depth, float* image) {
  for (uint32_t n = 0; n < N; ++n) {
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
    if (!(num_steps == 0 || offset + num_steps >= M)) {
      for (uint32_t s = 0; s < num_steps; ++s) {
        const uint32_t i = offset + s;
        const float alpha = 1.0f - expf(-sigmas[i] * deltas[2 * i]);
        const float w = alpha * T;
        r = fmaf(w, rgbs[3 * i], r);
        g = fmaf(w, rgbs[3 * i + 1], g);
        b = fmaf(w, rgbs[3 * i + 2], b);
        t += deltas[2 * i + 1];
        d = fmaf(w, t, d);
        ws += w;
        T *= 1.0f - alpha;
      }
    }
    weights_sum[index] = ws;
    depth[index] = d;
    image[index * 3] = r, image[index * 3 + 1] = g, image[index * 3 + 2] = b;
  }
}

void ref_composite_train_bwd(const float* grad_ws, const float* grad_image, const float* sigmas, const float* rgbs,
                             const float* deltas, const int32_t* rays, const float* weights_sum, const float* image,
                             uint32_t M, uint32_t N, float* grad_sigmas, float* grad_rgbs) {
  for (uint32_t n = 0; n < N; ++n) {
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps >= M) continue;
    const float* gi = grad_image + 3 * index;
    const float rf = image[3 * index], gf = image[3 * index + 1], bf = image[3 * index + 2], wsf = weights_sum[index];
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
    for (uint32_t s = 0; s < num_steps; ++s) {
      const uint32_t i = offset + s;
      const float alpha = 1.0f - expf(-sigmas[i] * deltas[2 * i]);
      const float w = alpha * T;
      r = fmaf(w, rgbs[3 * i], r);
      g = fmaf(w, rgbs[3 * i + 1], g);
      b = fmaf(w, rgbs[3 * i + 2], b);
      ws += w;
      T *= 1.0f - alpha;
      grad_rgbs[3 * i] = gi[0] * w;
      grad_rgbs[3 * i + 1] = gi[1] * w;
      grad_rgbs[3 * i + 2] = gi[2] * w;
      grad_sigmas[i] = deltas[2 * i] * (gi[0] * (T * rgbs[3 * i] - (rf - r)) + gi[1] * (T * rgbs[3 * i + 1] - (gf - g)) +
                                        gi[2] * (T * rgbs[3 * i + 2] - (bf - b)) + grad_ws[index] * (T - (wsf - ws)));
    }
  }
}

/* ------------------------------------------------------------------ inference compositing + compaction */
void ref_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t,
                        const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                        float* image) {
  for (uint32_t n = 0; n < n_alive; ++n) {
    const int index = rays_alive[n];
    float t = rays_t[n];
    float ws = weights_sum[index], d = depth[index];
    float r = image[3 * index], g = image[3 * index + 1], b = image[3 * index + 2];
    uint32_t step = 0;
    while (step < n_step) {
      const uint32_t i = n * n_step + step;
      if (deltas[2 * i] == 0) break;
      const float alpha = 1.0f - expf(-sigmas[i] * deltas[2 * i]);
      const float T = 1 - ws;
      const float w = alpha * T;
      ws += w;
      t += deltas[2 * i + 1];
      d = fmaf(w, t, d);
      r = fmaf(w, rgbs[3 * i], r);
      g = fmaf(w, rgbs[3 * i + 1], g);
      b = fmaf(w, rgbs[3 * i + 2], b);
      if (T < 1e-4) break;
      ++step;
    }
    rays_t[n] = step < n_step ? -1.0f : t;
    weights_sum[index] = ws;
    depth[index] = d;
    image[3 * index] = r, image[3 * index + 1] = g, image[3 * index + 2] = b;
  }
}

void ref_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t,
                      const float* rays_t_old, int32_t* alive_counter) {
  for (uint32_t n = 0; n < n_alive; ++n) {
    if (rays_t_old[n] >= 0) {
      const int32_t i = alive_counter[0]++;
      rays_alive[i] = rays_alive_old[n];
      rays_t[i] = rays_t_old[n];
    }
  }
}
