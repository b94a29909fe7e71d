// K5: ray-state kernels of the IDR ray tracer (sphere tracing from both ends with back-off line search,
// 100-sample sign search, secant refinement, minimal-SDF search for the mask loss).
//
// Follows model/ray_tracing.py of the reference statement by statement:
//   sphere_tracing :98-187, ray_sampler :189-249, secant :251-268, minimal_sdf_points :270-298,
//   forward :26-95.  Every arithmetic step uses the same fp32 operation sequence as the reference's
//   elementwise torch ops (separately rounded mul / add / div, no FMA contraction), so with the same SDF
//   values the hit/miss masks, distances and points are bit-identical.
//
// Instead of boolean-mask indexing + host syncs, each phase COMPACTS the rays that need an SDF
// evaluation into a dense point list (warp-aggregated atomic append) and remembers the list slot per
// ray; the SDF (MLP tiles + hash encode, or any user callable) is then evaluated "in place" on that
// list and the next phase gathers its value by slot.  Loop exits of the reference (`break` when no ray is
// unfinished) become device-side gates on a counter, so a whole trace can run without host round trips.
#include "common.cuh"

namespace idrk {

struct RayState {
    const float* cam;      // [B,3]
    const float* dirs;     // [N,3]
    int n, num_pixels;
    float *t0, *t1, *cur_s, *cur_e, *nxt_s, *nxt_e, *ps, *pe, *min_dis, *max_dis;
    unsigned char *unf_s, *unf_e;
    int *slot_s, *slot_e;
};

__device__ __forceinline__ float3 ray_point(const RayState& S, int i, float t) {
    const float* c = S.cam + 3 * (i / S.num_pixels);
    const float* d = S.dirs + 3 * (long long)i;
    return make_float3(__fadd_rn(c[0], __fmul_rn(t, d[0])), __fadd_rn(c[1], __fmul_rn(t, d[1])),
                       __fadd_rn(c[2], __fmul_rn(t, d[2])));
}
__device__ __forceinline__ void store3(float* p, long long i, float3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// warp-aggregated append: every lane asks for `want` (0..2) consecutive slots
__device__ __forceinline__ int warp_append(int* counter, int want) {
    const unsigned lane = threadIdx.x & 31;
    int incl = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - want;
}

__global__ void rt_init_kernel(RayState S, const float* __restrict__ t_sph, const unsigned char* __restrict__ hit,
                               float* __restrict__ pts, int* __restrict__ counter) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < S.n;
    const bool h = in && hit[i];
    const int slot = warp_append(counter, h ? 2 : 0);
    if (!in) return;
    float tn = 0.f, tf = 0.f;
    float3 a = make_float3(0.f, 0.f, 0.f), b = a;
    if (h) { tn = t_sph[2 * i]; tf = t_sph[2 * i + 1]; a = ray_point(S, i, tn); b = ray_point(S, i, tf); }
    S.t0[i] = tn; S.t1[i] = tf; S.min_dis[i] = tn; S.max_dis[i] = tf;
    store3(S.ps, i, a); store3(S.pe, i, b);
    S.unf_s[i] = h; S.unf_e[i] = h;
    S.nxt_s[i] = 0.f; S.nxt_e[i] = 0.f; S.cur_s[i] = 0.f; S.cur_e[i] = 0.f;
    S.slot_s[i] = h ? slot : -1; S.slot_e[i] = h ? slot + 1 : -1;
    if (h) { store3(pts, slot, a); store3(pts, slot + 1, b); }
}

// gather_mode: 0 none, 1 replace-all (nxt = slot>=0 ? val : 0), 2 update (slot>=0 -> nxt = val)
__device__ __forceinline__ void gather_vals(const RayState& S, int i, const float* __restrict__ vals, int mode) {
    if (mode == 0) return;
    const int a = S.slot_s[i], b = S.slot_e[i];
    if (mode == 1) { S.nxt_s[i] = a >= 0 ? vals[a] : 0.f; S.nxt_e[i] = b >= 0 ? vals[b] : 0.f; }
    else { if (a >= 0) S.nxt_s[i] = vals[a]; if (b >= 0) S.nxt_e[i] = vals[b]; }
}

// loop top (ray_tracing.py:131-142): current sdf, threshold, unfinished masks, count of unfinished rays
__global__ void rt_top_kernel(RayState S, const float* __restrict__ vals, int gather_mode, float thr, int* __restrict__ n_unf) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool any = false;
    if (i < S.n) {
        gather_vals(S, i, vals, gather_mode);
        bool us = S.unf_s[i], ue = S.unf_e[i];
        float cs = us ? S.nxt_s[i] : 0.f, ce = ue ? S.nxt_e[i] : 0.f;
        if (cs <= thr) cs = 0.f;
        if (ce <= thr) ce = 0.f;
        us = us && (cs > thr); ue = ue && (ce > thr);
        S.cur_s[i] = cs; S.cur_e[i] = ce; S.unf_s[i] = us; S.unf_e[i] = ue;
        any = us || ue;
    }
    const unsigned m = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_unf, __popc(m));
}

// make a step (:150-162) and list the unfinished end points for evaluation
__global__ void rt_step_kernel(RayState S, const int* __restrict__ gate, float* __restrict__ pts, int* __restrict__ counter) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    if (*gate == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < S.n;
    bool us = false, ue = false;
    float3 a, b;
    if (in) {
        const float t0 = __fadd_rn(S.t0[i], S.cur_s[i]);
        const float t1 = __fsub_rn(S.t1[i], S.cur_e[i]);
        S.t0[i] = t0; S.t1[i] = t1;
        a = ray_point(S, i, t0); b = ray_point(S, i, t1);
        store3(S.ps, i, a); store3(S.pe, i, b);
        us = S.unf_s[i]; ue = S.unf_e[i];
    }
    const int slot = warp_append(counter, (us ? 1 : 0) + (ue ? 1 : 0));
    if (!in) return;
    int sa = -1, sb = -1;
    if (us) { sa = slot; store3(pts, sa, a); }
    if (ue) { sb = slot + (us ? 1 : 0); store3(pts, sb, b); }
    S.slot_s[i] = sa; S.slot_e[i] = sb;
}

// one back-off iteration of the line search (:167-183)
__global__ void rt_linesearch_kernel(RayState S, const int* __restrict__ gate, const float* __restrict__ vals, int gather_mode,
                                     float factor, float* __restrict__ pts, int* __restrict__ counter) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    if (*gate == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < S.n;
    bool bs = false, be = false;
    float3 a, b;
    if (in) {
        gather_vals(S, i, vals, gather_mode);
        bs = S.nxt_s[i] < 0.f; be = S.nxt_e[i] < 0.f;
        if (bs) { const float t = __fsub_rn(S.t0[i], __fmul_rn(factor, S.cur_s[i])); S.t0[i] = t; a = ray_point(S, i, t); store3(S.ps, i, a); }
        if (be) { const float t = __fadd_rn(S.t1[i], __fmul_rn(factor, S.cur_e[i])); S.t1[i] = t; b = ray_point(S, i, t); store3(S.pe, i, b); }
    }
    const int slot = warp_append(counter, (bs ? 1 : 0) + (be ? 1 : 0));
    if (!in) return;
    int sa = -1, sb = -1;
    if (bs) { sa = slot; store3(pts, sa, a); }
    if (be) { sb = slot + (bs ? 1 : 0); store3(pts, sb, b); }
    S.slot_s[i] = sa; S.slot_e[i] = sb;
}

// The whole back-off line search of one iteration as ONE evaluation (:167-183).  The reference re-evaluates a ray whose
// new SDF is negative up to `line_step_iters` times, each time a little further back: t_k = t_{k-1} -+ f_k * cur_sdf,
// stopping at the first non-negative value.  The candidate positions depend only on the state BEFORE the search, and an
// SDF value does not depend on which other points share its batch, so all candidates of all offending rays are listed
// at once (n_ls consecutive slots per ray end), evaluated together, and resolved per ray in order: the first k whose
// value is not negative (else the last) becomes the ray's t / point / next sdf - exactly the state the sequential search
// leaves.  One SDF query and two launches per iteration instead of n_ls of each.
struct LsFactors { float f[8]; };

__global__ void rt_linesearch_points_kernel(RayState S, const int* __restrict__ gate, const float* __restrict__ vals,
                                            LsFactors F, int n_ls, float* __restrict__ pts, int* __restrict__ counter) {
    pdl_wait();
    pdl_trigger();
    if (*gate == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < S.n;
    bool bs = false, be = false;
    if (in) {
        gather_vals(S, i, vals, 1);                        // values of the step's evaluation
        bs = S.nxt_s[i] < 0.f; be = S.nxt_e[i] < 0.f;
    }
    const int slot = warp_append(counter, (bs ? n_ls : 0) + (be ? n_ls : 0));
    if (!in) return;
    int sa = -1, sb = -1;
    if (bs) {
        sa = slot;
        float t = S.t0[i];
        const float cur = S.cur_s[i];
        for (int k = 0; k < n_ls; ++k) { t = __fsub_rn(t, __fmul_rn(F.f[k], cur)); store3(pts, sa + k, ray_point(S, i, t)); }
    }
    if (be) {
        sb = slot + (bs ? n_ls : 0);
        float t = S.t1[i];
        const float cur = S.cur_e[i];
        for (int k = 0; k < n_ls; ++k) { t = __fadd_rn(t, __fmul_rn(F.f[k], cur)); store3(pts, sb + k, ray_point(S, i, t)); }
    }
    S.slot_s[i] = sa; S.slot_e[i] = sb;
}

__global__ void rt_linesearch_resolve_kernel(RayState S, const int* __restrict__ gate, const float* __restrict__ vals,
                                             LsFactors F, int n_ls) {
    pdl_wait();
    pdl_trigger();
    if (*gate == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    const int sa = S.slot_s[i], sb = S.slot_e[i];
    if (sa >= 0) {
        float t = S.t0[i], v = 0.f;
        const float cur = S.cur_s[i];
        for (int k = 0; k < n_ls; ++k) {
            t = __fsub_rn(t, __fmul_rn(F.f[k], cur));
            v = vals[sa + k];
            if (!(v < 0.f)) break;                         // no longer offending: the sequential search stops here
        }
        S.t0[i] = t; S.nxt_s[i] = v; store3(S.ps, i, ray_point(S, i, t));
    }
    if (sb >= 0) {
        float t = S.t1[i], v = 0.f;
        const float cur = S.cur_e[i];
        for (int k = 0; k < n_ls; ++k) {
            t = __fadd_rn(t, __fmul_rn(F.f[k], cur));
            v = vals[sb + k];
            if (!(v < 0.f)) break;
        }
        S.t1[i] = t; S.nxt_e[i] = v; store3(S.pe, i, ray_point(S, i, t));
    }
    S.slot_s[i] = -1; S.slot_e[i] = -1;
}

// end of an iteration (:185-186)
__global__ void rt_end_kernel(RayState S, const int* __restrict__ gate, const float* __restrict__ vals, int gather_mode) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    if (*gate == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    gather_vals(S, i, vals, gather_mode);
    const bool ok = S.t0[i] < S.t1[i];
    S.unf_s[i] = S.unf_s[i] && ok;
    S.unf_e[i] = S.unf_e[i] && ok;
}

// Tail of a sphere-tracing iteration as ONE launch: resolve the line search (above), close the iteration (rt_end_kernel)
// and open the next one (rt_top_kernel without a gather) - three per-ray passes over the same state with no dependence
// between rays; only the count of unfinished rays (the next iteration's gate) crosses threads, through one atomic per warp.
__global__ void rt_iter_tail_kernel(RayState S, const int* __restrict__ gate, const float* __restrict__ vals, LsFactors F, int n_ls,
                                    float thr, int* __restrict__ n_unf) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = *gate != 0;
    bool any = false;
    if (i < S.n) {
        float t0 = S.t0[i], t1 = S.t1[i];
        float ns = S.nxt_s[i], ne = S.nxt_e[i];
        bool us = S.unf_s[i], ue = S.unf_e[i];
        if (live) {
            const int sa = S.slot_s[i], sb = S.slot_e[i];
            if (sa >= 0) {
                float t = t0, v = 0.f;
                const float cur = S.cur_s[i];
                for (int k = 0; k < n_ls; ++k) {
                    t = __fsub_rn(t, __fmul_rn(F.f[k], cur));
                    v = vals[sa + k];
                    if (!(v < 0.f)) break;
                }
                t0 = t; ns = v; S.t0[i] = t; S.nxt_s[i] = v; store3(S.ps, i, ray_point(S, i, t));
            }
            if (sb >= 0) {
                float t = t1, v = 0.f;
                const float cur = S.cur_e[i];
                for (int k = 0; k < n_ls; ++k) {
                    t = __fadd_rn(t, __fmul_rn(F.f[k], cur));
                    v = vals[sb + k];
                    if (!(v < 0.f)) break;
                }
                t1 = t; ne = v; S.t1[i] = t; S.nxt_e[i] = v; store3(S.pe, i, ray_point(S, i, t));
            }
            S.slot_s[i] = -1; S.slot_e[i] = -1;
            const bool ok = t0 < t1;                       // rt_end_kernel
            us = us && ok; ue = ue && ok;
        }
        float cs = us ? ns : 0.f, ce = ue ? ne : 0.f;      // rt_top_kernel, gather_mode 0
        if (cs <= thr) cs = 0.f;
        if (ce <= thr) ce = 0.f;
        us = us && (cs > thr); ue = ue && (ce > thr);
        S.cur_s[i] = cs; S.cur_e[i] = ce; S.unf_s[i] = us; S.unf_e[i] = ue;
        any = us || ue;
    }
    const unsigned m = __ballot_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_unf, __popc(m));
}

// after sphere tracing (:39-42): network mask and the list of rays handed to the sampler
__global__ void rt_select_sampler_kernel(RayState S, unsigned char* __restrict__ net_mask, int* __restrict__ ray_of_slot,
                                         int* __restrict__ counter) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < S.n;
    const bool sel = in && S.unf_s[i];
    const int slot = warp_append(counter, sel ? 1 : 0);
    if (!in) return;
    net_mask[i] = S.t0[i] < S.t1[i];
    S.slot_s[i] = sel ? slot : -1;
    if (sel) ray_of_slot[slot] = i;
}

// sample points of slots [slot0, slot0 + n_slots): z = tmin + lin[j] * (tmax - tmin)   (:197-200)
// n_dev (nullable): device-side number of valid slots; the chunk then covers slots [slot0, min(slot0 + n_slots, *n_dev))
__device__ __forceinline__ int chunk_slots(int slot0, int n_slots, const int* n_dev) {
    if (n_dev == nullptr) return n_slots;
    const int left = *n_dev - slot0;
    return left < 0 ? 0 : (left < n_slots ? left : n_slots);
}

__global__ void rt_chunk_counts_kernel(const int* __restrict__ n_dev, int per_chunk, int mult, int n_chunks, int* __restrict__ out) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chunks) out[c] = chunk_slots(c * per_chunk, per_chunk, n_dev) * mult;
}

__global__ void rt_sampler_points_kernel(RayState S, const int* __restrict__ ray_of_slot, int slot0, int n_slots, int n_steps,
                                         const float* __restrict__ lin, float* __restrict__ pts, const int* __restrict__ n_dev) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    n_slots = chunk_slots(slot0, n_slots, n_dev);
    const long long total = (long long)n_slots * n_steps;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(k / n_steps), j = (int)(k - (long long)s * n_steps);
        const int i = ray_of_slot[slot0 + s];
        const float lo = S.t0[i], hi = S.t1[i];
        const float z = __fadd_rn(lo, __fmul_rn(lin[j], __fsub_rn(hi, lo)));
        store3(pts, k, ray_point(S, i, z));
    }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// per sampler ray (one warp): first sign change, fallback minimum, secant bracket (:212-247)
__global__ void rt_sampler_resolve_kernel(RayState S, const int* __restrict__ ray_of_slot, int n_slots, int n_steps,
                                          const float* __restrict__ lin, const float* __restrict__ vals,
                                          const unsigned char* __restrict__ object_mask, int training,
                                          unsigned char* __restrict__ net_mask, float* __restrict__ z_lo, float* __restrict__ z_hi,
                                          float* __restrict__ s_lo, float* __restrict__ s_hi, int* __restrict__ sec_slots,
                                          int* __restrict__ sec_counter, const int* __restrict__ n_dev) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    n_slots = chunk_slots(0, n_slots, n_dev);
    if (s >= n_slots) return;
    const int i = ray_of_slot[s];
    const float* v = vals + (long long)s * n_steps;
    // argmin_j sign(v_j) * (n_steps - j), first occurrence;  argmin_j v_j, first occurrence
    float best_key = INFINITY, best_val = INFINITY;
    int best_kj = 0x7fffffff, best_vj = 0x7fffffff;
    for (int j = lane; j < n_steps; j += 32) {
        const float vj = v[j];
        const float key = __fmul_rn(sgn(vj), (float)(n_steps - j));
        if (key < best_key) { best_key = key; best_kj = j; }
        if (vj < best_val) { best_val = vj; best_vj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, best_key, o); const int oj = __shfl_xor_sync(0xffffffffu, best_kj, o);
        if (ok < best_key || (ok == best_key && oj < best_kj)) { best_key = ok; best_kj = oj; }
        const float ov = __shfl_xor_sync(0xffffffffu, best_val, o); const int ovj = __shfl_xor_sync(0xffffffffu, best_vj, o);
        if (ov < best_val || (ov == best_val && ovj < best_vj)) { best_val = ov; best_vj = ovj; }
    }
    if (lane != 0) return;
    const float lo = S.t0[i], hi = S.t1[i];
    const float span = __fsub_rn(hi, lo);
    const int first = best_kj;
    const float v_first = v[first];
    const bool net_surf = v_first < 0.f;
    const bool true_surf = object_mask[i] != 0;
    int pick = first;
    if (!(true_surf && net_surf)) pick = best_vj;
    const float z_pick = __fadd_rn(lo, __fmul_rn(lin[pick], span));
    const bool secant = training ? (net_surf && true_surf) : net_surf;
    net_mask[i] = net_surf;
    if (secant) {
        const int prev = first > 0 ? first - 1 : n_steps - 1;       // index -1 wraps (reference :239)
        z_hi[s] = __fadd_rn(lo, __fmul_rn(lin[first], span)); s_hi[s] = v_first;
        z_lo[s] = __fadd_rn(lo, __fmul_rn(lin[prev], span));  s_lo[s] = v[prev];
        sec_slots[atomicAdd(sec_counter, 1)] = s;
    } else {
        S.t0[i] = z_pick;
        store3(S.ps, i, ray_point(S, i, z_pick));
    }
}

__device__ __forceinline__ float secant_pred(float s_lo, float s_hi, float z_lo, float z_hi) {
    // - sdf_low * (z_high - z_low) / (sdf_high - sdf_low) + z_low      (:253)
    return __fadd_rn(__fdiv_rn(__fmul_rn(-s_lo, __fsub_rn(z_hi, z_lo)), __fsub_rn(s_hi, s_lo)), z_lo);
}

// secant iteration: mode 0 = emit first prediction points, 1 = consume sdf_mid + emit next, 2 = consume + write result,
// 3 = write the initial prediction (n_secant_steps == 0)
__global__ void rt_secant_kernel(RayState S, const int* __restrict__ ray_of_slot, const int* __restrict__ sec_slots, int n_sec,
                                 int mode, const float* __restrict__ vals, float* __restrict__ z_lo, float* __restrict__ z_hi,
                                 float* __restrict__ s_lo, float* __restrict__ s_hi, float* __restrict__ pts,
                                 const int* __restrict__ n_dev) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    n_sec = chunk_slots(0, n_sec, n_dev);
    if (k >= n_sec) return;
    const int s = sec_slots[k];
    const int i = ray_of_slot[s];
    float zl = z_lo[s], zh = z_hi[s], sl = s_lo[s], sh = s_hi[s];
    if (mode == 1 || mode == 2) {
        const float zp = secant_pred(sl, sh, zl, zh);
        const float sm = vals[k];
        if (sm > 0.f) { zl = zp; sl = sm; }
        if (sm < 0.f) { zh = zp; sh = sm; }
        z_lo[s] = zl; z_hi[s] = zh; s_lo[s] = sl; s_hi[s] = sh;
    }
    const float zp = secant_pred(sl, sh, zl, zh);
    const float3 p = ray_point(S, i, zp);
    if (mode >= 2) { S.t0[i] = zp; store3(S.ps, i, p); }
    else store3(pts, k, p);
}

// training-only tail (:71-92): rays outside the sphere, list of rays for the minimal-SDF search
__global__ void rt_select_minsdf_kernel(RayState S, const unsigned char* __restrict__ net_mask, const unsigned char* __restrict__ object_mask,
                                        const unsigned char* __restrict__ hit, const unsigned char* __restrict__ sampler_mask,
                                        int* __restrict__ ray_of_slot, int* __restrict__ counter) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < S.n;
    bool sel = false;
    if (in) {
        const bool nm = net_mask[i], om = object_mask[i], sm = sampler_mask[i], h = hit[i];
        const bool in_mask = !nm && om && !sm;
        const bool out_mask = !om && !sm;
        if ((in_mask || out_mask) && !h) {       // project the camera centre onto the ray (:78-82)
            const float* c = S.cam + 3 * (i / S.num_pixels);
            const float* d = S.dirs + 3 * (long long)i;
            const float dot = __fadd_rn(__fadd_rn(__fmul_rn(d[0], c[0]), __fmul_rn(d[1], c[1])), __fmul_rn(d[2], c[2]));
            const float t = -dot;
            S.t0[i] = t;
            store3(S.ps, i, ray_point(S, i, t));
        }
        sel = (in_mask || out_mask) && h;
        if (sel && nm && out_mask) S.min_dis[i] = S.t0[i];     // :87
    }
    const int slot = warp_append(counter, sel ? 1 : 0);
    if (!in) return;
    S.slot_s[i] = sel ? slot : -1;
    if (sel) ray_of_slot[slot] = i;
}

// steps = u[j] * (max_dis - min_dis) + min_dis  (:277-286)
__global__ void rt_minsdf_points_kernel(RayState S, const int* __restrict__ ray_of_slot, int slot0, int n_slots, int n_steps,
                                        const float* __restrict__ u, float* __restrict__ pts, const int* __restrict__ n_dev) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    n_slots = chunk_slots(slot0, n_slots, n_dev);
    const long long total = (long long)n_slots * n_steps;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(k / n_steps), j = (int)(k - (long long)s * n_steps);
        const int i = ray_of_slot[slot0 + s];
        const float lo = S.min_dis[i], hi = S.max_dis[i];
        const float z = __fadd_rn(__fmul_rn(u[j], __fsub_rn(hi, lo)), lo);
        store3(pts, k, ray_point(S, i, z));
    }
}

__global__ void rt_minsdf_resolve_kernel(RayState S, const int* __restrict__ ray_of_slot, int n_slots, int n_steps,
                                         const float* __restrict__ u, const float* __restrict__ vals, const int* __restrict__ n_dev) {
    pdl_wait();                 // programmatic dependent launch: everything below may touch the producer's data
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    n_slots = chunk_slots(0, n_slots, n_dev);
    if (s >= n_slots) return;
    const int i = ray_of_slot[s];
    const float* v = vals + (long long)s * n_steps;
    float best = INFINITY; int bj = 0x7fffffff;
    for (int j = lane; j < n_steps; j += 32) { const float vj = v[j]; if (vj < best) { best = vj; bj = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov < best || (ov == best && oj < bj)) { best = ov; bj = oj; }
    }
    if (lane != 0) return;
    if (bj >= n_steps) bj = 0;
    const float lo = S.min_dis[i], hi = S.max_dis[i];
    const float z = __fadd_rn(__fmul_rn(u[bj], __fsub_rn(hi, lo)), lo);
    S.t0[i] = z;
    store3(S.ps, i, ray_point(S, i, z));
}

static int fill_state(const idrk_ray_state_t* h, RayState& S) {
    if (!h || h->n_rays < 0 || h->num_pixels < 1) return IDRK_E_ARG;
    const void* need[] = {h->cam_loc, h->ray_dirs, h->t0, h->t1, h->cur_s, h->cur_e, h->nxt_s, h->nxt_e, h->ps, h->pe,
                          h->min_dis, h->max_dis, h->unf_s, h->unf_e, h->slot_s, h->slot_e};
    for (const void* p : need) if (!p) return IDRK_E_ARG;
    S.cam = h->cam_loc; S.dirs = h->ray_dirs; S.n = h->n_rays; S.num_pixels = h->num_pixels;
    S.t0 = h->t0; S.t1 = h->t1; S.cur_s = h->cur_s; S.cur_e = h->cur_e; S.nxt_s = h->nxt_s; S.nxt_e = h->nxt_e;
    S.ps = h->ps; S.pe = h->pe; S.min_dis = h->min_dis; S.max_dis = h->max_dis;
    S.unf_s = h->unf_s; S.unf_e = h->unf_e; S.slot_s = h->slot_s; S.slot_e = h->slot_e;
    return 0;
}

}  // namespace idrk

using namespace idrk;

#define RT_PRELUDE()                                   \
    RayState S;                                        \
    { int rc = fill_state(h_state, S); if (rc) return rc; } \
    if (S.n == 0) return 0;                            \
    const int threads = 256;                           \
    const int blocks = (S.n + threads - 1) / threads;  \
    cudaStream_t st = (cudaStream_t)stream;

extern "C" int idrk_rt_init(const idrk_ray_state_t* h_state, const float* t_sph, const uint8_t* hit, float* pts, int32_t* counter,
                            void* stream) {
    RT_PRELUDE();
    if (!t_sph || !hit || !pts || !counter) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_init_kernel, dim3(blocks), dim3(threads), 0, st, S, t_sph, hit, pts, counter));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_top(const idrk_ray_state_t* h_state, const float* vals, int32_t gather_mode, float sdf_threshold,
                           int32_t* n_unfinished, void* stream) {
    RT_PRELUDE();
    if (!n_unfinished || (gather_mode && !vals)) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_top_kernel, dim3(blocks), dim3(threads), 0, st, S, vals, gather_mode, sdf_threshold, n_unfinished));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_step(const idrk_ray_state_t* h_state, const int32_t* gate, float* pts, int32_t* counter, void* stream) {
    RT_PRELUDE();
    if (!gate || !pts || !counter) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_step_kernel, dim3(blocks), dim3(threads), 0, st, S, gate, pts, counter));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_linesearch(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, int32_t gather_mode,
                                  float factor, float* pts, int32_t* counter, void* stream) {
    RT_PRELUDE();
    if (!gate || !pts || !counter || (gather_mode && !vals)) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_linesearch_kernel, dim3(blocks), dim3(threads), 0, st, S, gate, vals, gather_mode, factor, pts, counter));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_linesearch_points(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals,
                                         const float* h_factors, int32_t n_ls, float* pts, int32_t* counter, void* stream) {
    RT_PRELUDE();
    if (!gate || !vals || !h_factors || !pts || !counter || n_ls < 1 || n_ls > 8) return IDRK_E_ARG;
    LsFactors F;
    for (int k = 0; k < 8; ++k) F.f[k] = k < n_ls ? h_factors[k] : 0.f;
    IDRK_CUDA_TRY(launch_k(rt_linesearch_points_kernel, dim3(blocks), dim3(threads), 0, st, S, gate, vals, F, (int)n_ls, pts, counter));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_linesearch_resolve(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals,
                                          const float* h_factors, int32_t n_ls, void* stream) {
    RT_PRELUDE();
    if (!gate || !vals || !h_factors || n_ls < 1 || n_ls > 8) return IDRK_E_ARG;
    LsFactors F;
    for (int k = 0; k < 8; ++k) F.f[k] = k < n_ls ? h_factors[k] : 0.f;
    IDRK_CUDA_TRY(launch_k(rt_linesearch_resolve_kernel, dim3(blocks), dim3(threads), 0, st, S, gate, vals, F, (int)n_ls));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_iter_tail(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, const float* h_factors,
                                 int32_t n_ls, float sdf_threshold, int32_t* n_unfinished, void* stream) {
    RT_PRELUDE();
    if (!gate || !vals || !h_factors || !n_unfinished || n_ls < 1 || n_ls > 8) return IDRK_E_ARG;
    LsFactors F;
    for (int k = 0; k < 8; ++k) F.f[k] = k < n_ls ? h_factors[k] : 0.f;
    IDRK_CUDA_TRY(launch_k(rt_iter_tail_kernel, dim3(blocks), dim3(threads), 0, st, S, gate, vals, F, (int)n_ls, sdf_threshold, n_unfinished));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_end(const idrk_ray_state_t* h_state, const int32_t* gate, const float* vals, int32_t gather_mode, void* stream) {
    RT_PRELUDE();
    if (!gate || (gather_mode && !vals)) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_end_kernel, dim3(blocks), dim3(threads), 0, st, S, gate, vals, gather_mode));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_select_sampler(const idrk_ray_state_t* h_state, uint8_t* net_mask, int32_t* ray_of_slot, int32_t* counter,
                                      void* stream) {
    RT_PRELUDE();
    if (!net_mask || !ray_of_slot || !counter) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_select_sampler_kernel, dim3(blocks), dim3(threads), 0, st, S, net_mask, ray_of_slot, counter));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_sampler_points(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t slot0, int32_t n_slots,
                                      int32_t n_steps, const float* lin, float* pts, const int32_t* n_dev, void* stream) {
    RT_PRELUDE();
    (void)blocks;
    if (!ray_of_slot || !lin || !pts || n_slots < 0 || n_steps < 1) return IDRK_E_ARG;
    if (n_slots == 0) return 0;
    long long b = ((long long)n_slots * n_steps + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    IDRK_CUDA_TRY(launch_k(rt_sampler_points_kernel, dim3((int)b), dim3(threads), 0, st, S, ray_of_slot, slot0, n_slots, n_steps, lin, pts, n_dev));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_sampler_resolve(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t n_slots, int32_t n_steps,
                                       const float* lin, const float* vals, const uint8_t* object_mask, int32_t training,
                                       uint8_t* net_mask, float* z_lo, float* z_hi, float* s_lo, float* s_hi,
                                       int32_t* sec_slots, int32_t* sec_counter, const int32_t* n_dev, void* stream) {
    RT_PRELUDE();
    (void)blocks;
    if (!ray_of_slot || !lin || !vals || !object_mask || !net_mask || !z_lo || !z_hi || !s_lo || !s_hi || !sec_slots || !sec_counter)
        return IDRK_E_ARG;
    if (n_slots <= 0) return 0;
    IDRK_CUDA_TRY(launch_k(rt_sampler_resolve_kernel, dim3((n_slots + 7) / 8), dim3(256), 0, st, S, ray_of_slot, n_slots, n_steps, lin, vals, object_mask, training,
                                                                net_mask, z_lo, z_hi, s_lo, s_hi, sec_slots, sec_counter, n_dev));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_secant(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, const int32_t* sec_slots, int32_t n_sec,
                              int32_t mode, const float* vals, float* z_lo, float* z_hi, float* s_lo, float* s_hi, float* pts,
                              const int32_t* n_dev, void* stream) {
    RT_PRELUDE();
    (void)blocks;
    if (!ray_of_slot || !sec_slots || !z_lo || !z_hi || !s_lo || !s_hi || ((mode == 1 || mode == 2) && !vals) || (mode < 2 && !pts)) return IDRK_E_ARG;
    if (n_sec <= 0) return 0;
    IDRK_CUDA_TRY(launch_k(rt_secant_kernel, dim3((n_sec + threads - 1) / threads), dim3(threads), 0, st, S, ray_of_slot, sec_slots, n_sec, mode, vals, z_lo, z_hi,
                                                                       s_lo, s_hi, pts, n_dev));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_select_minsdf(const idrk_ray_state_t* h_state, const uint8_t* net_mask, const uint8_t* object_mask,
                                     const uint8_t* hit, const uint8_t* sampler_mask, int32_t* ray_of_slot, int32_t* counter,
                                     void* stream) {
    RT_PRELUDE();
    if (!net_mask || !object_mask || !hit || !sampler_mask || !ray_of_slot || !counter) return IDRK_E_ARG;
    IDRK_CUDA_TRY(launch_k(rt_select_minsdf_kernel, dim3(blocks), dim3(threads), 0, st, S, net_mask, object_mask, hit, sampler_mask, ray_of_slot, counter));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_minsdf_points(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t slot0, int32_t n_slots,
                                     int32_t n_steps, const float* u, float* pts, const int32_t* n_dev, void* stream) {
    RT_PRELUDE();
    (void)blocks;
    if (!ray_of_slot || !u || !pts || n_slots < 0 || n_steps < 1) return IDRK_E_ARG;
    if (n_slots == 0) return 0;
    long long b = ((long long)n_slots * n_steps + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    IDRK_CUDA_TRY(launch_k(rt_minsdf_points_kernel, dim3((int)b), dim3(threads), 0, st, S, ray_of_slot, slot0, n_slots, n_steps, u, pts, n_dev));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_minsdf_resolve(const idrk_ray_state_t* h_state, const int32_t* ray_of_slot, int32_t n_slots, int32_t n_steps,
                                      const float* u, const float* vals, const int32_t* n_dev, void* stream) {
    RT_PRELUDE();
    (void)blocks;
    if (!ray_of_slot || !u || !vals) return IDRK_E_ARG;
    if (n_slots <= 0) return 0;
    IDRK_CUDA_TRY(launch_k(rt_minsdf_resolve_kernel, dim3((n_slots + 7) / 8), dim3(256), 0, st, S, ray_of_slot, n_slots, n_steps, u, vals, n_dev));
    IDRK_LAUNCH_CHECK();
    return 0;
}

extern "C" int idrk_rt_chunk_counts(const int32_t* n_dev, int32_t per_chunk, int32_t mult, int32_t n_chunks, int32_t* out, void* stream) {
    if (!n_dev || !out || per_chunk < 1 || mult < 1 || n_chunks < 0) return IDRK_E_ARG;
    if (n_chunks == 0) return 0;
    IDRK_CUDA_TRY(launch_k(rt_chunk_counts_kernel, dim3((n_chunks + 127) / 128), dim3(128), 0, (cudaStream_t)stream, n_dev, per_chunk, mult, n_chunks, out));
    IDRK_LAUNCH_CHECK();
    return 0;
}
