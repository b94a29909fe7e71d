"""RayTracing with the reference's constructor and call signature (model/ray_tracing.py:5-95):

    RayTracing(object_bounding_sphere, sdf_threshold, line_search_step, line_step_iters,
               sphere_tracing_iters, n_steps, n_secant_steps)
    .forward(sdf, cam_loc, object_mask, ray_directions) -> (points, network_object_mask, dists)

The per-ray state machine runs in csrc/ray_tracing.cu; between its phases the SDF is evaluated on
compacted point lists.  Two ways to evaluate:

* `sdf` is `ImplicitNetwork.sdf` of this package with a device-count capable encoder: the whole
  sphere-tracing loop (up to 44 SDF evaluations) is enqueued without a single host sync - kernels take
  the list length from device memory; two syncs remain (sampler / min-SDF list sizes);
* any other callable (e.g. an analytic SDF in the parity tests, or the reference's lambda): the list
  length is read back after each compaction and `sdf(points[:n])` is called like the reference does.

Extensions (keyword-only, default to the reference behaviour): `min_sdf_steps` injects the U(0,1)
vector of ray_tracing.py:277 and `sphere_intersections=(t, hit)` the result of
rend_util.get_sphere_intersection, so parity runs can share host-side randomness / inputs.
"""
import ctypes

import torch
import torch.nn as nn

from .. import kernels as K
from .._lib import RayStateDesc, check, lib, ptr, stream_ptr
from ..utils import rend_util

import os as _os

# Points per SDF call in the sampler / min-SDF sweeps.  Round 1 used 32768 (activations of a chunk stay in L2); with the
# fp16-pair pipeline the sweep's layers run at the same speed from HBM, and fewer, larger chunks mean fewer launches, no
# gated-off chunk launches and no short tail chunk: measured whole step 6.48 ms (32768) / 6.35 (65536) / 6.27-6.33 (262144,
# one chunk for 2048 rays) / 6.28-6.30 (2^19, 2^20); cfg3 at 65536 rays 149 -> 141 ms, cfg4 165 -> 155 ms.
_SDF_CHUNK_POINTS = int(_os.environ.get("IDRK_SDF_CHUNK", "262144"))


class _Evaluator:
    """Evaluates the SDF on the first `count` rows of a point buffer."""

    def __init__(self, sdf):
        owner = getattr(sdf, "__self__", None)
        self.fast = (owner is not None and hasattr(owner, "sdf_compacted") and owner.supports_device_count()
                     and getattr(sdf, "__func__", None) is getattr(type(owner), "sdf", None))
        self.owner = owner
        self.sdf = sdf
        self.calls = 0
        self.points = 0

    def on_device_count(self, pts, cap, counter, out):
        """Sphere-tracing phases: list length lives in `counter` (device int32[1])."""
        self.calls += 1
        if self.fast:
            self.owner.sdf_compacted(pts, cap, counter, out)
            return
        n = int(counter.item())
        self.points += n
        if n > 0:
            out[:n] = self.sdf(pts[:n]).reshape(-1)

    def on_host_count(self, pts, n, out):
        self.calls += 1
        self.points += n
        if n == 0:
            return
        if self.fast:
            self.owner.sdf_compacted(pts, n, None, out)
        else:
            out[:n] = self.sdf(pts[:n]).reshape(-1)


class _TraceState:
    """Per-(B, P, device) persistent ray-state buffers (so the launch sequence can live in a CUDA graph)."""

    def __init__(self, B, P, n_it, n_ls, dev):
        N = B * P
        self.B, self.P, self.N, self.dev = B, P, N, dev
        self.t_sph = torch.empty((N, 2), device=dev, dtype=torch.float32)
        self.hit_u8 = torch.empty(N, device=dev, dtype=torch.uint8)
        self.cam = torch.empty((B, 3), device=dev, dtype=torch.float32)
        self.dirs = torch.empty((N, 3), device=dev, dtype=torch.float32)
        fbuf = torch.empty(16 * N, device=dev, dtype=torch.float32)
        (self.t0, self.t1, self.cur_s, self.cur_e, self.nxt_s, self.nxt_e, self.min_dis,
         self.max_dis) = (fbuf[i * N:(i + 1) * N] for i in range(8))
        self.ps = fbuf[8 * N:11 * N].view(N, 3)
        self.pe = fbuf[11 * N:14 * N].view(N, 3)
        self._fbuf = fbuf
        bbuf = torch.empty(3 * N, device=dev, dtype=torch.uint8)
        self.unf_s, self.unf_e, self.net_mask = bbuf[:N], bbuf[N:2 * N], bbuf[2 * N:]
        ibuf = torch.empty(3 * N, device=dev, dtype=torch.int32)
        self.slot_s, self.slot_e, self.ray_of_slot = ibuf[:N], ibuf[N:2 * N], ibuf[2 * N:]
        self.counters = torch.zeros(8 + (n_it + 1) * (n_ls + 3), device=dev, dtype=torch.int32)
        self.cidx = 0
        self.cap = 2 * N * max(1, n_ls)          # both ends of every ray x the line search's candidates per end
        self.pts = torch.empty((self.cap, 3), device=dev, dtype=torch.float32)
        self.vals = torch.empty(self.cap, device=dev, dtype=torch.float32)
        st = RayStateDesc()
        st.cam_loc, st.ray_dirs, st.n_rays, st.num_pixels = self.cam.data_ptr(), self.dirs.data_ptr(), N, P
        for name in ("t0", "t1", "cur_s", "cur_e", "nxt_s", "nxt_e", "ps", "pe", "min_dis", "max_dis", "unf_s", "unf_e",
                     "slot_s", "slot_e"):
            setattr(st, name, getattr(self, name).data_ptr())
        self.desc = st
        self.graphs = {}
        self.retired = []
        self.warm = set()
        self._tail = None

    def tail(self, ns):
        """Static buffers of the device-driven sampler / secant / min-SDF phases (allocated on first use)."""
        if self._tail is None or self._tail["ns"] != ns:
            N, dev = self.N, self.dev
            rpc = max(1, min(_SDF_CHUNK_POINTS // ns, N))        # never more rows than the batch can need
            n_chunks = (N + rpc - 1) // rpc
            zbuf = torch.empty(4 * N, device=dev, dtype=torch.float32)
            self._tail = {
                "ns": ns, "rpc": rpc, "n_chunks": n_chunks,
                "obj_u8": torch.empty(N, device=dev, dtype=torch.uint8),
                "sampler_mask": torch.empty(N, device=dev, dtype=torch.uint8),
                "u": torch.empty(ns, device=dev, dtype=torch.float32),
                "lin": torch.linspace(0, 1, steps=ns).to(dev),
                "chunk_cnt": torch.zeros(n_chunks, device=dev, dtype=torch.int32),
                "big_vals": torch.empty(N * ns, device=dev, dtype=torch.float32),
                "cpts": torch.empty((rpc * ns, 3), device=dev, dtype=torch.float32),
                "z": [zbuf[i * N:(i + 1) * N] for i in range(4)],
                "sec_slots": torch.empty(N, device=dev, dtype=torch.int32),
                "spts": torch.empty((N, 3), device=dev, dtype=torch.float32),
                "svals": torch.empty(N, device=dev, dtype=torch.float32),
                "counts": None,
            }
        return self._tail

    def new_counter(self):
        c = self.counters[self.cidx:self.cidx + 1]
        self.cidx += 1
        return c


class RayTracing(nn.Module):
    def __init__(self, object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=1,
                 sphere_tracing_iters=10, n_steps=100, n_secant_steps=8):
        super().__init__()
        self.object_bounding_sphere = object_bounding_sphere
        self.sdf_threshold = sdf_threshold
        self.sphere_tracing_iters = sphere_tracing_iters
        self.line_step_iters = line_step_iters
        self.line_search_step = line_search_step
        self.n_steps = n_steps
        self.n_secant_steps = n_secant_steps
        self._stats = {}
        self._pending_counts = None
        self.injected_min_sdf_steps = None      # parity runs: the U(0,1) vector of reference :277
        self.sample_generator = None            # optional torch.Generator for that draw (per-rank in data-parallel runs)
        self.use_cuda_graph = False             # replay the sphere-tracing launch sequence from a CUDA graph
        self._states = {}

    # ------------------------------------------------------------------------------------------
    def _sphere_trace(self, T, ev):
        """Enqueues the whole bidirectional sphere-tracing loop (reference :98-187): a fixed launch sequence whose
        data-dependent exits are device-side gates, so it can be captured in a CUDA graph."""
        L = lib()
        S = ctypes.byref(T.desc)
        sp = stream_ptr()
        n_ls, n_it = int(self.line_step_iters), int(self.sphere_tracing_iters)
        T.counters.zero_()
        T.cidx = 0
        c = T.new_counter()
        check(L.idrk_rt_init(S, ptr(T.t_sph), ptr(T.hit_u8), ptr(T.pts), ptr(c), sp), "idrk_rt_init")
        ev.on_device_count(T.pts, 2 * T.N, c, T.vals)
        gate = T.new_counter()
        check(L.idrk_rt_top(S, ptr(T.vals), 1, float(self.sdf_threshold), ptr(gate), sp), "idrk_rt_top")
        if n_ls > 8:
            raise ValueError("line_step_iters > 8 is not supported")
        factors = (ctypes.c_float * max(n_ls, 1))(*[(1 - self.line_search_step) / (2 ** k) for k in range(n_ls)])
        for it in range(n_it):
            c = T.new_counter()
            check(L.idrk_rt_step(S, ptr(gate), ptr(T.pts), ptr(c), sp), "idrk_rt_step")
            ev.on_device_count(T.pts, 2 * T.N, c, T.vals)
            if n_ls > 0:
                # the whole back-off search in ONE evaluation: every candidate of every offending ray end (csrc/ray_tracing.cu)
                c = T.new_counter()
                check(L.idrk_rt_linesearch_points(S, ptr(gate), ptr(T.vals), factors, n_ls, ptr(T.pts), ptr(c), sp),
                      "idrk_rt_linesearch_points")
                ev.on_device_count(T.pts, T.cap, c, T.vals)
                # resolve the search, close the iteration and open the next one (its gate) in one launch
                nxt = T.new_counter()
                check(L.idrk_rt_iter_tail(S, ptr(gate), ptr(T.vals), factors, n_ls, float(self.sdf_threshold), ptr(nxt), sp),
                      "idrk_rt_iter_tail")
                gate = nxt
                continue
            check(L.idrk_rt_end(S, ptr(gate), ptr(T.vals), 1, sp), "idrk_rt_end")
            gate = T.new_counter()
            check(L.idrk_rt_top(S, None, 0, float(self.sdf_threshold), ptr(gate), sp), "idrk_rt_top")

    def _tail_device(self, T, ev):
        """Sampler + secant (+ min-SDF when training) with every list length kept on the device
        (reference :41-59, :71-92, :189-298): a fixed launch sequence, no host reads."""
        L = lib()
        S = ctypes.byref(T.desc)
        sp = stream_ptr()
        ns = int(self.n_steps)
        t = T.tail(ns)
        N, rpc, n_chunks = T.N, t["rpc"], t["n_chunks"]
        z_lo, z_hi, s_lo, s_hi = t["z"]
        c_samp = T.new_counter()
        check(L.idrk_rt_select_sampler(S, ptr(T.net_mask), ptr(T.ray_of_slot), ptr(c_samp), sp), "idrk_rt_select_sampler")
        t["sampler_mask"].copy_(T.unf_s)
        check(L.idrk_rt_chunk_counts(ptr(c_samp), rpc, ns, n_chunks, ptr(t["chunk_cnt"]), sp), "idrk_rt_chunk_counts")
        for c in range(n_chunks):
            check(L.idrk_rt_sampler_points(S, ptr(T.ray_of_slot), c * rpc, rpc, ns, ptr(t["lin"]), ptr(t["cpts"]),
                                           ptr(c_samp), sp), "idrk_rt_sampler_points")
            ev.on_device_count(t["cpts"], rpc * ns, t["chunk_cnt"][c:c + 1], t["big_vals"][c * rpc * ns:])
        c_sec = T.new_counter()
        check(L.idrk_rt_sampler_resolve(S, ptr(T.ray_of_slot), N, ns, ptr(t["lin"]), ptr(t["big_vals"]), ptr(t["obj_u8"]),
                                        int(self.training), ptr(T.net_mask), ptr(z_lo), ptr(z_hi), ptr(s_lo), ptr(s_hi),
                                        ptr(t["sec_slots"]), ptr(c_sec), ptr(c_samp), sp), "idrk_rt_sampler_resolve")

        def secant(mode):
            check(L.idrk_rt_secant(S, ptr(T.ray_of_slot), ptr(t["sec_slots"]), N, mode, ptr(t["svals"]), ptr(z_lo),
                                   ptr(z_hi), ptr(s_lo), ptr(s_hi), ptr(t["spts"]), ptr(c_sec), sp), "idrk_rt_secant")
        n_secant = int(self.n_secant_steps)
        if n_secant == 0:
            secant(3)
        else:
            secant(0)
            for i in range(n_secant):
                ev.on_device_count(t["spts"], N, c_sec, t["svals"])
                secant(1 if i < n_secant - 1 else 2)
        c_min = None
        if self.training:
            c_min = T.new_counter()
            check(L.idrk_rt_select_minsdf(S, ptr(T.net_mask), ptr(t["obj_u8"]), ptr(T.hit_u8), ptr(t["sampler_mask"]),
                                          ptr(T.ray_of_slot), ptr(c_min), sp), "idrk_rt_select_minsdf")
            check(L.idrk_rt_chunk_counts(ptr(c_min), rpc, ns, n_chunks, ptr(t["chunk_cnt"]), sp), "idrk_rt_chunk_counts")
            for c in range(n_chunks):
                check(L.idrk_rt_minsdf_points(S, ptr(T.ray_of_slot), c * rpc, rpc, ns, ptr(t["u"]), ptr(t["cpts"]),
                                              ptr(c_min), sp), "idrk_rt_minsdf_points")
                ev.on_device_count(t["cpts"], rpc * ns, t["chunk_cnt"][c:c + 1], t["big_vals"][c * rpc * ns:])
            check(L.idrk_rt_minsdf_resolve(S, ptr(T.ray_of_slot), N, ns, ptr(t["u"]), ptr(t["big_vals"]), ptr(c_min), sp),
                  "idrk_rt_minsdf_resolve")
        t["counts"] = (c_samp, c_sec, c_min)

    def _trace_device(self, T, ev):
        self._sphere_trace(T, ev)
        self._tail_device(T, ev)

    @property
    def last_stats(self):
        """Tracer statistics of the last call (reads the device counters lazily: one host sync when accessed)."""
        st = dict(self._stats)
        pend = self._pending_counts
        if pend is not None:
            c_samp, c_sec, c_min = pend
            st.update({"n_sampler": int(c_samp.item()), "n_secant": int(c_sec.item())})
            if c_min is not None:
                st["n_minsdf"] = int(c_min.item())
        return st

    def _state(self, B, P, dev):
        key = (B, P, str(dev), int(self.sphere_tracing_iters), int(self.line_step_iters))
        T = self._states.get(key)
        if T is None:
            T = self._states[key] = _TraceState(B, P, int(self.sphere_tracing_iters), int(self.line_step_iters), dev)
        return T

    def forward(self, sdf, cam_loc, object_mask, ray_directions, *, min_sdf_steps=None, sphere_intersections=None):
        K.require_cuda(ray_directions, "ray_directions")
        L = lib()
        dev = ray_directions.device
        B, P, _ = ray_directions.shape
        N = B * P
        ev = _Evaluator(sdf)
        with torch.no_grad():
            if sphere_intersections is None:
                t_sph, hit = rend_util.get_sphere_intersection(cam_loc, ray_directions, r=self.object_bounding_sphere)
            else:
                t_sph, hit = sphere_intersections
            T = self._state(B, P, dev)
            T.t_sph.copy_(t_sph.reshape(N, 2))
            T.hit_u8.copy_(hit.reshape(N))
            T.cam.copy_(cam_loc.reshape(B, 3))
            T.dirs.copy_(ray_directions.reshape(N, 3))
            hit_u8 = T.hit_u8
            obj_u8 = object_mask.reshape(N).to(torch.uint8).contiguous()
            t0, ps, unf_s, net_mask, ray_of_slot = T.t0, T.ps, T.unf_s, T.net_mask, T.ray_of_slot
            S = ctypes.byref(T.desc)
            sp = stream_ptr()
            new_counter = T.new_counter

            if ev.fast:
                # ---- fully device-driven trace: fixed launch sequence, optionally replayed from a CUDA graph ----
                t = T.tail(int(self.n_steps))
                t["obj_u8"].copy_(obj_u8)
                if self.training:
                    if min_sdf_steps is None:
                        min_sdf_steps = self.injected_min_sdf_steps
                    if min_sdf_steps is None:       # drawn on the host generator like the reference (:277)
                        min_sdf_steps = torch.empty(int(self.n_steps)).uniform_(0.0, 1.0, generator=self.sample_generator)
                    t["u"].copy_(min_sdf_steps.to(dev, non_blocking=True).float())
                key = bool(self.training)
                # a graph is only valid for the scratch buffers it was captured against (kernels.SCRATCH_GENERATION)
                if key in T.graphs and T.graphs[key][3] != K.SCRATCH_GENERATION[0]:
                    T.retired.append(T.graphs.pop(key))
                    T.warm.discard(key)
                if not self.use_cuda_graph:
                    self._trace_device(T, ev)
                elif key not in T.warm:              # first call: eager (allocations, lazy attribute setup)
                    self._trace_device(T, ev)
                    T.warm.add(key)
                else:
                    if key not in T.graphs:
                        torch.cuda.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        ev.owner.refresh_inference_weights()
                        l0 = K._lib.LAUNCHES[0]
                        with torch.cuda.graph(graph):
                            ev.owner.refresh_inference_weights(force=True)   # weight folding is part of the graph
                            self._trace_device(T, ev)
                        K.note_graph_captured()
                        T.graphs[key] = (graph, K._lib.LAUNCHES[0] - l0, ev.calls, K.SCRATCH_GENERATION[0])
                    graph, n_launches, n_calls, _ = T.graphs[key]
                    graph.replay()
                    K._lib.LAUNCHES[0] += n_launches            # kernels launched by the replayed graph
                    ev.calls = n_calls
                self._stats = {"sdf_calls": ev.calls, "fast_path": True, "cuda_graph": bool(self.use_cuda_graph)}
                self._pending_counts = t["counts"]
                return ps.clone(), net_mask.bool().clone(), t0.clone()

            # ---- generic callable: host reads the list lengths and calls sdf(points[:n]) -------------------------
            self._sphere_trace(T, ev)

            # ---- sampler + secant for the non-convergent rays (:41-59, :189-268) ------------------------
            c_samp = new_counter()
            check(L.idrk_rt_select_sampler(S, ptr(net_mask), ptr(ray_of_slot), ptr(c_samp), sp), "idrk_rt_select_sampler")
            sampler_mask = unf_s.clone()
            n_samp = int(c_samp.item())
            n_sec = 0
            ns = int(self.n_steps)
            if n_samp > 0:
                lin = torch.linspace(0, 1, steps=ns).to(dev)
                big_vals = torch.empty(n_samp * ns, device=dev, dtype=torch.float32)
                rays_per_chunk = max(1, _SDF_CHUNK_POINTS // ns)
                cpts = torch.empty((min(n_samp, rays_per_chunk) * ns, 3), device=dev, dtype=torch.float32)
                for s0 in range(0, n_samp, rays_per_chunk):
                    m = min(rays_per_chunk, n_samp - s0)
                    check(L.idrk_rt_sampler_points(S, ptr(ray_of_slot), s0, m, ns, ptr(lin), ptr(cpts), None, sp),
                          "idrk_rt_sampler_points")
                    ev.on_host_count(cpts, m * ns, big_vals[s0 * ns:(s0 + m) * ns])
                zbuf = torch.empty(4 * n_samp, device=dev, dtype=torch.float32)
                z_lo, z_hi, s_lo, s_hi = (zbuf[i * n_samp:(i + 1) * n_samp] for i in range(4))
                sec_slots = torch.empty(n_samp, device=dev, dtype=torch.int32)
                c_sec = new_counter()
                check(L.idrk_rt_sampler_resolve(S, ptr(ray_of_slot), n_samp, ns, ptr(lin), ptr(big_vals), ptr(obj_u8),
                                                int(self.training), ptr(net_mask), ptr(z_lo), ptr(z_hi), ptr(s_lo),
                                                ptr(s_hi), ptr(sec_slots), ptr(c_sec), None, sp), "idrk_rt_sampler_resolve")
                n_sec = int(c_sec.item())
                if n_sec > 0:
                    spts = torch.empty((n_sec, 3), device=dev, dtype=torch.float32)
                    svals = torch.empty(n_sec, device=dev, dtype=torch.float32)
                    nsec_steps = int(self.n_secant_steps)

                    def secant(mode):
                        check(L.idrk_rt_secant(S, ptr(ray_of_slot), ptr(sec_slots), n_sec, mode, ptr(svals), ptr(z_lo),
                                               ptr(z_hi), ptr(s_lo), ptr(s_hi), ptr(spts), None, sp), "idrk_rt_secant")
                    if nsec_steps == 0:
                        secant(3)
                    else:
                        secant(0)
                        for i in range(nsec_steps):
                            ev.on_host_count(spts, n_sec, svals)
                            secant(1 if i < nsec_steps - 1 else 2)

            self._stats = {"sdf_calls": ev.calls, "n_sampler": n_samp, "n_secant": n_sec, "fast_path": ev.fast}
            self._pending_counts = None
            net_mask_b = net_mask.bool()
            if not self.training:
                return ps.clone(), net_mask_b.clone(), t0.clone()

            # ---- training only: rays that miss, minimal-SDF points for the mask loss (:71-92, :270-298) --
            c_min = new_counter()
            check(L.idrk_rt_select_minsdf(S, ptr(net_mask), ptr(obj_u8), ptr(hit_u8), ptr(sampler_mask), ptr(ray_of_slot),
                                          ptr(c_min), sp), "idrk_rt_select_minsdf")
            n_min = int(c_min.item())
            if n_min > 0:
                if min_sdf_steps is None:
                    min_sdf_steps = self.injected_min_sdf_steps
                if min_sdf_steps is None:       # drawn on the host generator like the reference (:277)
                    u = torch.empty(ns).uniform_(0.0, 1.0, generator=self.sample_generator).to(dev)
                else:
                    u = min_sdf_steps.to(dev).float().contiguous()
                big_vals = torch.empty(n_min * ns, device=dev, dtype=torch.float32)
                rays_per_chunk = max(1, _SDF_CHUNK_POINTS // ns)
                cpts = torch.empty((min(n_min, rays_per_chunk) * ns, 3), device=dev, dtype=torch.float32)
                for s0 in range(0, n_min, rays_per_chunk):
                    m = min(rays_per_chunk, n_min - s0)
                    check(L.idrk_rt_minsdf_points(S, ptr(ray_of_slot), s0, m, ns, ptr(u), ptr(cpts), None, sp),
                          "idrk_rt_minsdf_points")
                    ev.on_host_count(cpts, m * ns, big_vals[s0 * ns:(s0 + m) * ns])
                check(L.idrk_rt_minsdf_resolve(S, ptr(ray_of_slot), n_min, ns, ptr(u), ptr(big_vals), None, sp),
                      "idrk_rt_minsdf_resolve")
            self._stats.update({"sdf_calls": ev.calls, "n_minsdf": n_min})
            return ps.clone(), net_mask_b.clone(), t0.clone()
