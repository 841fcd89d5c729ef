"""Graph-captured nerfacto training step: the B200-native runtime of the hot path.

`Trainer` in nerfacto.py drives the drop-in modules through torch.autograd (dozens of small framework ops and
Python dispatch per step).  `NerfactoStep` runs the SAME arithmetic — the step of
nerfstudio/engine/trainer.py:486-530 over models/nerfacto.py:298-391 — as a fixed sequence of our own kernel
launches on preallocated buffers (forward, losses, hand-written backward, fused Adam), captures that sequence in
a CUDA graph and replays it: one graph launch per optimisation step, no tracing compiler, no autograd tape, no
allocation.  Parameters stay the model's own nn.Parameters (views into one flat buffer), so state_dicts and
evaluation through the module API are unaffected.

Per-step scalars that change (Adam bias corrections of the three parameter groups, proposal-weight anneal exponent, lr
schedules) live in a tiny device buffer refreshed with one small copy before each replay; the stratified draws come from a
Philox state kept on the device.

Shape of the captured step:
  * single process — ONE graph whose independent pieces are parallel branches (forked streams during capture): the
    proposal fields' backward + their Adam, the main grid's position gradient, the gradient memset + TMA weight packing,
    the per-ray-constant columns of the colour head's input, the embedding-row gradients;
  * one process per GPU — four graph pieces with the collectives between them: the field segment's gradients are
    reduce-scattered, every rank runs Adam on its 1/N slice, the parameter all-gather lands under the next step's proposal
    sampling (`sharded_update=False`: two all-reduces over the flat buffer, replicated Adam).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict, List, Optional

import torch
from torch import Tensor

from . import functional as F
from . import lib
from .lib import B2nMlpGrad, call, ptr, stream
from .nerfacto import NerfactoModel
from .optim import FlatAdam

NULL = C.c_void_p(0)


def _off(t: Tensor, elems: int) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() + 4 * elems)


class _Net:
    """One hash-grid + MLP network bound to its parameters / gradient views."""

    def __init__(self, encoding, mlp):
        self.grid = encoding.grid
        self.table = encoding.hash_table
        self.spec = mlp.spec
        self.weights = [l.weight for l in mlp.layers]
        self.biases = [l.bias for l in mlp.layers]

    def structs(self):
        m = self.spec.struct(self.weights, self.biases)
        g = B2nMlpGrad()
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            g.dw[i], g.db[i] = ptr(w.grad).value, ptr(b.grad).value  # NULL when there is no optimiser (render engine)
        return m, g


class NerfactoStep:
    HYPER_SLOTS = 64

    def __init__(self, model: NerfactoModel, n_rays: int, lr: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-15,
                 lr_schedule: Optional[Callable[[int], float]] = None, allreduce=None, use_graph: bool = True,
                 always_update_proposals: bool = False, mlp_backend: str = "auto",
                 fused_proposals: bool = True, eval_mode: bool = False, camera_lr: float = 1e-3,
                 camera_lr_schedule: Optional[Callable[[int], float]] = None, fused_tail: bool = True,
                 concurrent_backward: bool = True, sharded_update: bool = True) -> None:
        cfg = model.config
        if cfg.implementation != "torch":
            raise NotImplementedError("the captured step is built on the torch-mode (parity) networks")
        if cfg.use_same_proposal_network or cfg.num_proposal_iterations != 2:
            raise NotImplementedError("captured step: two separate proposal networks (the nerfacto default)")
        self.model, self.cfg, self.R = model, cfg, n_rays
        # eval_mode (render_engine.NerfactoRender): forward only, deterministic samplers, no optimiser / gradient buffers
        self.eval_mode = eval_mode
        self.optim = None if eval_mode else FlatAdam(model, lr=lr, betas=betas, eps=eps, lr_schedule=lr_schedule)
        self.allreduce, self.use_graph = allreduce, use_graph
        # the proposal networks' segment of the flat buffer (FlatAdam groups parameters by Model.get_param_groups()).
        # When it is the tail of the buffer the gradient all-reduce is split there (field segment summed while the
        # proposal backward still runs); any other layout gets ONE all-reduce after the whole backward — never a
        # collective over memory the proposal backward is still accumulating into.
        # camera optimiser (nerfacto's default: models/nerfacto.py:131, applied in get_outputs :300-301).  Mode SO3xR3 trains in
        # the captured step: pose_apply at the head of the forward, position gradients of all three levels back to the
        # rays, pose_apply_bwd + the regulariser at the tail, its own Adam group (lr 1e-3: method_configs.py:114-117).
        cam = getattr(model, "camera_optimizer", None)
        cam_mode = getattr(getattr(cam, "config", None), "mode", "off") if cam is not None else "off"
        if cam_mode not in ("off", "SO3xR3"):
            raise NotImplementedError(f"captured step: camera optimizer mode {cam_mode!r} (only off / SO3xR3)")
        self.camopt = cam if (cam_mode == "SO3xR3" and not eval_mode) else None
        self.camera_lr, self.camera_lr_schedule = camera_lr, camera_lr_schedule
        # gradient all-reduce layout: the flat buffer is [... | segments that are complete only after the proposal / pose
        # backward].  When "camera_opt" and "proposal_networks" form the tail, the head (the field) is summed over NVLink
        # while that backward still runs; any other layout gets ONE all-reduce after the whole backward.
        total = self.optim.flat.numel() if self.optim is not None else 0
        self.grad_split, self._cam_seg, self._prop_seg = None, None, None
        if self.optim is not None:
            self._prop_seg = self.optim.segment_of("proposal_networks")
            self._cam_seg = self.optim.segment_of("camera_opt") if self.camopt is not None else None
            late = sorted(sg for sg in (self._cam_seg, self._prop_seg) if sg is not None)
            if late and late[-1][1] == total and all(a[1] == b[0] for a, b in zip(late, late[1:])) and late[0][0] > 0 \
                    and (self._prop_seg is not None):
                self.grad_split = late[0][0]
        self._prop_steps = 0  # optimiser steps the proposal group has taken (its own bias-correction count)
        self.always_update = always_update_proposals
        dev = next(model.parameters()).device
        self.dev = dev
        self.S = list(cfg.num_proposal_samples_per_ray) + [cfg.num_nerf_samples_per_ray]
        self.props = [_Net(p.encoding, p.mlp_base[1]) for p in model.proposal_networks]
        fld = model.field
        self.base = _Net(fld.mlp_base.model[0], fld.mlp_base.model[1])
        self.head_spec = fld.mlp_head.spec
        self.head_w = [l.weight for l in fld.mlp_head.layers]
        self.head_b = [l.bias for l in fld.mlp_head.layers]
        self.emb = fld.embedding_appearance.embedding.weight if fld.embedding_appearance is not None else None
        self.geo, self.n_emb = fld.geo_feat_dim, (fld.appearance_embedding_dim if self.emb is not None else 0)
        specs = [p_.spec for p_ in self.props] + [self.base.spec, self.head_spec]
        if mlp_backend == "tc" and not all(F.mlp_tc_supported(sp) for sp in specs):
            raise NotImplementedError("a network of this model is outside the tensor-core MLP's shape range")

        def use_tc(sp) -> bool:
            if mlp_backend == "simt" or not F.mlp_tc_supported(sp):
                return False
            # auto: tensor cores pay off once a layer is >= 32 wide; the 10->16->1 proposal nets stay on FFMA
            return mlp_backend == "tc" or max(sp.out_dims) >= 32

        self.tc_net = {id(sp): use_tc(sp) for sp in specs}
        self._tc_ws: Dict[int, Tensor] = {}
        self._live_ws: Dict[int, Tensor] = {}
        self.compact_live = True  # proposal backward visits only the samples whose density gradient is non-zero
        self.tma_weights = True  # stage the tensor-core MLP weights by TMA from a per-step packed image
        self.fused_props = fused_proposals and all(F.density_field_supported(p_.grid, p_.spec) for p_ in self.props)
        self.tc = use_tc(self.head_spec)
        self.contraction = fld.spatial_distortion is not None
        self.aabb = fld.aabb.flatten().tolist()
        self.avg = float(cfg.average_init_density)
        self.spacing = "uniform" if cfg.proposal_initial_sampler == "uniform" else "piecewise"
        self.bg = cfg.background_color
        R = n_rays
        f32 = dict(device=dev, dtype=torch.float32)
        # ---- static inputs
        # one device blob [camera indices (int64) | origins | directions | gt rgb]: a packed pinned host batch
        # (pack_batch) reaches the GPU with a single async copy; the four tensors are views of it
        self.inputs = torch.zeros(11 * R, **f32)
        self.cams = self.inputs[: 2 * R].view(torch.int64)
        self.origins_in = self.inputs[2 * R: 5 * R].view(R, 3)      # what the data manager delivers
        self.directions_in = self.inputs[5 * R: 8 * R].view(R, 3)
        self.gt = self.inputs[8 * R: 11 * R].view(R, 3)
        if self.camopt is not None:  # rays after the per-camera pose correction (what every kernel downstream reads)
            self.origins, self.directions = torch.zeros(R, 3, **f32), torch.zeros(R, 3, **f32)
            self.d_rays = torch.zeros(2, R, 3, **f32)           # d loss / d (corrected origins | directions)
            self.d_x2 = torch.zeros(R * cfg.num_nerf_samples_per_ray, 3, **f32)
            from .cameras.camera_optimizers import _frozen_mask

            self.cam_frozen = _frozen_mask(self.camopt)
            self.cam_pose = self.camopt.pose_adjustment
        else:
            self.origins, self.directions = self.origins_in, self.directions_in
        # NearFarCollider (scene_colliders.py:169-191): the near plane is reset to 0 outside training
        self.nears = torch.full((R,), 0.0 if eval_mode else float(cfg.near_plane), **f32)
        self.fars = torch.full((R,), float(cfg.far_plane), **f32)
        # [lr/bc1, 1/sqrt(bc2), grad_scale, anneal | the same three for the proposal group's own step count, pad |
        #  the same three with the camera optimiser's learning rate, pad]
        self.hyper = torch.zeros(12, **f32)
        # pinned staging ring for the per-step scalars: a slot is rewritten only after the async H2D copy that read it
        # has completed (the CPU may run many steps ahead of the stream)
        self._hyper_ring = torch.zeros(self.HYPER_SLOTS, 12, dtype=torch.float32).pin_memory()
        self._hyper_np = self._hyper_ring.numpy()  # same memory; one vectorised write per step
        self._hyper_events: List[Optional[torch.cuda.Event]] = [None] * self.HYPER_SLOTS
        self.lin0 = torch.linspace(0.0, 1.0, self.S[0] + 1).to(dev)
        self.u_base = [torch.linspace(0.0, 1.0 - 1.0 / (s + 1), s + 1).to(dev) for s in self.S[1:]]
        # ---- per-level buffers
        self.sb, self.eb, self.x, self.sel, self.enc, self.h, self.hid, self.dens, self.w = [], [], [], [], [], [], [], [], []
        self.d_w, self.d_dens, self.d_h, self.d_enc = [], [], [], []
        for lvl, S in enumerate(self.S):
            N = R * S
            net = self.props[lvl] if lvl < 2 else self.base
            width_out = net.spec.out_dims[-1]
            self.sb.append(torch.zeros(R, S + 1, **f32)), self.eb.append(torch.zeros(R, S + 1, **f32))
            self.x.append(torch.zeros(N, 3, **f32)), self.sel.append(torch.zeros(N, device=dev, dtype=torch.uint8))
            self.enc.append(torch.zeros(N, net.grid.out_dim, **f32))
            self.h.append(torch.zeros(N, width_out, **f32))
            self.hid.append(torch.zeros(max(net.spec.hidden_width, 1) * N, **f32))
            self.dens.append(torch.zeros(R, S, **f32)), self.w.append(torch.zeros(R, S, **f32))
            self.d_w.append(torch.zeros(R, S, **f32)), self.d_dens.append(torch.zeros(R, S, **f32))
            self.d_h.append(torch.zeros(N, width_out, **f32)), self.d_enc.append(torch.zeros(N, net.grid.out_dim, **f32))
        N2 = R * self.S[2]
        self.n_sh = 16
        self.head_in_w = self.n_sh + self.geo + self.n_emb
        self.hin_stride = (self.head_in_w + 3) // 4 * 4 if self.tc else self.head_in_w  # padded rows: 128-bit loads
        self.sh = torch.zeros(R, self.n_sh, **f32)
        self.hin, self.d_hin = torch.zeros(N2, self.hin_stride, **f32), torch.zeros(N2, self.hin_stride, **f32)
        self.hid_head = torch.zeros(self.head_spec.hidden_width * N2, **f32)
        self.rgb, self.d_rgb = torch.zeros(N2, 3, **f32), torch.zeros(N2, 3, **f32)
        self.d_hpre = torch.zeros(N2, **f32)
        self.d_w_dist = torch.zeros(R, self.S[2], **f32)
        self.rgb_out, self.d_rgb_out = torch.zeros(R, 3, **f32), torch.zeros(R, 3, **f32)
        self.acc, self.depth_exp, self.depth_med = torch.zeros(R, **f32), torch.zeros(R, **f32), torch.zeros(R, **f32)
        self.prop_depth = [torch.zeros(R, **f32) for _ in range(2)]  # median depth of the proposal levels (eval outputs)
        self.emb_mean = torch.zeros(1, max(self.n_emb, 1), **f32)    # eval: mean appearance embedding (or zeros)
        self.rows = [torch.zeros(R, **f32) for _ in range(4)]  # per-ray loss terms: interlevel 0/1, distortion, rgb
        # the per-ray middle of the step (weights, renderers, losses, their backward) in one launch (csrc/ray_tail.cu);
        # False = one operator per launch (the A/B reference of tests/test_gpu_engine.py)
        self.fused_tail = fused_tail and not eval_mode
        # Independent pieces of the backward on forked streams (joined before the pose backward / Adam; inside the captured
        # graph they become parallel branches): the proposal fields' backward only needs the ray tail's d_density, and the
        # main grid's position gradient (gather) is independent of its table scatter (atomics).  In the sparse-gradient
        # regime the proposal backward is a few latency-bound rounds on a fraction of the SMs — it hides completely.
        # (i) is not used when collectives run between the pieces of the step (the field all-reduce overlaps the proposal
        # backward there); the other branches are joined inside the first piece.
        self.concurrent = concurrent_backward and not eval_mode
        self._side = [torch.cuda.Stream(), torch.cuda.Stream()] if self.concurrent else []
        self._fork = False        # grad zeroing / weight packing / position-gradient branches (joined inside _body)
        self._fork_props = False  # proposal backward as a branch (joined in _body_props; single-process step only)
        self._joins = []
        self._props_stepped = False  # the proposal group was already stepped inside the forked branch of this step
        self._head_static_join = None
        self._prologue_join = None
        self._pack_late = False   # sharded update: the weight images are packed after the parameter all-gather has landed
        self._h_ag = None         # pending all-gather of the field parameters (sharded update)
        self.sharded_update = sharded_update
        self._shard = None        # (chunk, my_begin, my_end, tail_begin) of the field segment while the sharded path is active
        self.losses = torch.zeros(5, **f32)  # rgb, interlevel, distortion, total, camera-optimiser regulariser
        self.jitter_all = torch.zeros(3, R, 1, **f32)  # stratified draws of the three sampling levels, one launch per step
        self.jitter = list(self.jitter_all.unbind(0))
        # Philox state {seed, draw counter, block tickets} on the device: every replay of the captured step draws afresh
        rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0x7FFFFFFFFFFFFFFF  # ranks draw independent streams
        self.rng_state = torch.tensor([seed, 0, 0], dtype=torch.int64, device=self.dev)
        self.step_count = 0
        self._graphs: Dict[bool, torch.cuda.CUDAGraph] = {}
        self._steps_since_update = 0
        self.fixed_jitter = None
        self.fixed_bins = None  # tests: [(spacing bins, euclidean bins)] per level replacing the samplers' outputs
        if self.bg not in ("last_sample", "white", "black"):
            raise NotImplementedError("captured step: background_color must be last_sample / white / black")

    # ------------------------------------------------------------------------------------------------
    def _tc_workspace(self, m, spec) -> Tensor:
        """Per-network device workspace for the packed weight images the tensor-core kernels stage with TMA."""
        ws = self._tc_ws.get(id(spec))
        if ws is None:
            nbytes = int(lib.load().b2n_mlp_tc_workspace_bytes(C.byref(m)))
            if nbytes <= 0:
                raise RuntimeError("b2n_mlp_tc_workspace_bytes failed")
            ws = torch.zeros(nbytes // 4 + 64, device=self.dev, dtype=torch.float32)
            off = (-ws.data_ptr()) % 128 // 4  # 128-byte aligned start (TMA source)
            ws = ws[off: off + nbytes // 4]
            self._tc_ws[id(spec)] = ws
        return ws

    def _pack_weights(self) -> None:
        """Once per step, before the first forward: operand images (hi/lo split, UMMA layout) of the current weights."""
        nets = [(self.base.structs()[0], self.base.spec), (self.head_spec.struct(self.head_w, self.head_b), self.head_spec)]
        nets += [(p_.structs()[0], p_.spec) for p_ in self.props]
        for m, spec in nets:
            if self.tc_net[id(spec)]:
                call("b2n_mlp_tc_pack", C.byref(m), ptr(self._tc_workspace(m, spec)), stream())

    def _mlp_fwd(self, m, x: Tensor, x_stride: int, n: int, y: Tensor, hidden: Tensor, spec=None) -> None:
        if self.tc_net[id(spec)]:
            call("b2n_mlp_tc_fwd_ws", C.byref(m), ptr(x), x_stride, n, ptr(y), ptr(hidden),
                 ptr(self._tc_workspace(m, spec)) if self.tma_weights else NULL, stream())
        else:
            call("b2n_mlp_fwd", C.byref(m), ptr(x), n, ptr(y), ptr(hidden), stream())

    def _mlp_bwd(self, m, g, x: Tensor, x_stride: int, y: Tensor, hidden: Tensor, dy: Tensor, n: int, dx: Tensor,
                 dx_stride: int, spec=None) -> None:
        if self.tc_net[id(spec)]:
            call("b2n_mlp_tc_bwd_ws", C.byref(m), C.byref(g), ptr(x), x_stride, ptr(y), ptr(hidden), ptr(dy), n, ptr(dx),
                 dx_stride, ptr(self._tc_workspace(m, spec)) if self.tma_weights else NULL, stream())
        else:
            call("b2n_mlp_bwd", C.byref(m), C.byref(g), ptr(x), ptr(y), ptr(hidden), ptr(dy), n, ptr(dx), stream())

    def _density_net_fwd(self, lvl: int, net: _Net, weights: bool = True) -> None:
        R, S = self.R, self.S[lvl]
        N = R * S
        eb = self.eb[lvl]
        box = lib.host_floats(self.aabb)
        if self.fused_props:
            m, _ = net.structs()
            call("b2n_density_field_fwd", C.byref(net.grid.c), C.byref(m), ptr(net.table), ptr(self.origins),
                 ptr(self.directions), ptr(eb), _off(eb, 1), S + 1, R, S, int(self.contraction), C.cast(box, C.c_void_p),
                 self.avg, ptr(self.dens[lvl]), stream())
            if weights:
                call("b2n_weights_fwd", ptr(eb), _off(eb, 1), S + 1, ptr(self.dens[lvl]), R, S, ptr(self.w[lvl]), stream())
            return
        call("b2n_positions_fwd", ptr(self.origins), ptr(self.directions), ptr(eb), _off(eb, 1), S + 1, R, S,
             int(self.contraction), C.cast(box, C.c_void_p), ptr(self.x[lvl]), ptr(self.sel[lvl], torch.uint8), stream())
        call("b2n_hashgrid_fwd", C.byref(net.grid.c), ptr(self.x[lvl]), ptr(net.table), N, ptr(self.enc[lvl]), NULL, stream())
        m, _ = net.structs()
        self._mlp_fwd(m, self.enc[lvl], self.enc[lvl].shape[1], N, self.h[lvl], self.hid[lvl], net.spec)
        call("b2n_density_act_fwd", ptr(self.h[lvl]), self.h[lvl].shape[1], ptr(self.sel[lvl], torch.uint8), N, self.avg,
             ptr(self.dens[lvl]), stream())
        if weights:
            call("b2n_weights_fwd", ptr(eb), _off(eb, 1), S + 1, ptr(self.dens[lvl]), R, S, ptr(self.w[lvl]), stream())

    def _density_net_bwd(self, lvl: int, net: _Net, d_hpre: Optional[Tensor]) -> None:
        """weights -> density -> pre-activation -> MLP -> encoding -> table, for a proposal level."""
        R, S = self.R, self.S[lvl]
        N = R * S
        eb = self.eb[lvl]
        if not self.fused_tail:  # (the fused tail has already turned d_w[lvl] into d_dens[lvl])
            call("b2n_weights_bwd", ptr(eb), _off(eb, 1), S + 1, ptr(self.dens[lvl]), ptr(self.d_w[lvl]), R, S,
                 ptr(self.d_dens[lvl]), stream())
        if self.fused_props:
            self.density_field_bwd_launch(lvl)
            return
        call("b2n_density_act_bwd", ptr(self.h[lvl]), 1, ptr(self.sel[lvl], torch.uint8), ptr(self.d_dens[lvl]), N, self.avg,
             ptr(self.d_h[lvl]), 1, stream())
        m, g = net.structs()
        self._mlp_bwd(m, g, self.enc[lvl], self.enc[lvl].shape[1], self.h[lvl], self.hid[lvl], self.d_h[lvl], N,
                      self.d_enc[lvl], self.d_enc[lvl].shape[1], net.spec)
        call("b2n_hashgrid_bwd", C.byref(net.grid.c), ptr(self.x[lvl]), ptr(net.table), ptr(self.d_enc[lvl]), N,
             ptr(net.table.grad), NULL, stream())

    def density_field_bwd_launch(self, lvl: int) -> None:
        """The fused proposal-field backward of level `lvl` on the current `d_dens[lvl]` (also used by bench.py to time
        the kernel in the dense-gradient regime: its cost depends on how many samples carry a gradient)."""
        net, R, S, eb = self.props[lvl], self.R, self.S[lvl], self.eb[lvl]
        m, g = net.structs()
        box = lib.host_floats(self.aabb)
        if lvl not in self._live_ws:
            self._live_ws[lvl] = torch.zeros(R * S + 4, device=self.dev, dtype=torch.int32)
        cam = self.camopt is not None
        call("b2n_density_field_bwd_rays", C.byref(net.grid.c), C.byref(m), C.byref(g), ptr(net.table), ptr(self.origins),
             ptr(self.directions), ptr(eb), _off(eb, 1), S + 1, R, S, int(self.contraction), C.cast(box, C.c_void_p),
             self.avg, ptr(self.d_dens[lvl]), ptr(net.table.grad),
             ptr(self._live_ws[lvl], torch.int32) if self.compact_live else NULL,
             ptr(self.d_rays[0]) if cam else NULL, ptr(self.d_rays[1]) if cam else NULL, stream())

    def _forward(self) -> None:
        """Rays in the static buffers -> samples of the three levels, densities, weights, colours, rendered outputs.
        Training: stratified samplers, per-camera appearance embedding.  Eval: deterministic samplers
        (ray_samplers.py:326-330), the mean embedding (or zeros), eval-mode compositing (renderers.py:225-231) and the
        proposal levels' median depths (models/nerfacto.py:346-347)."""
        self._forward_props()
        self._forward_main()

    def _forward_props(self) -> None:
        """Pose correction and the proposal sampling of both levels (reads only the proposal networks and the poses)."""
        R, S0, S1, S2 = self.R, *self.S
        st = stream
        ev = self.eval_mode
        prologue = self._prologue_join  # training step with forked streams: packing (and grad zeroing) already in flight
        if self.tma_weights and prologue is None and not self._pack_late:
            self._pack_weights()
        if prologue is not None and not self.fused_props:  # unfused proposal networks read their packed images right away
            torch.cuda.current_stream().wait_event(prologue)
            self._prologue_join = None
        if self.camopt is not None:  # CameraOptimizer.apply_to_raybundle (camera_optimizers.py:148-153)
            call("b2n_pose_apply_fwd", ptr(self.cam_pose), ptr(self.cams, torch.int64), ptr(self.cam_frozen, torch.uint8),
                 ptr(self.origins_in), ptr(self.directions_in), R, ptr(self.origins), ptr(self.directions), st())
        self._head_static_join = None
        if not self._pack_late:
            # SH of the (pose-corrected) directions and the per-ray constant columns of the colour head's input do not depend
            # on the field: a branch beside the proposal sampling (in the sharded multi-GPU step the embedding table is still
            # being all-gathered at this point: the branch starts at the head of _forward_main instead)
            self._fork_head_static()
        # ---------------- forward: proposal sampling
        call("b2n_spaced_sample", ptr(self.nears), ptr(self.fars), ptr(self.lin0), NULL if ev else ptr(self.jitter[0]), 0, R, S0,
             lib.SPACING[self.spacing], ptr(self.sb[0]), ptr(self.eb[0]), st())
        if self.fixed_bins is not None:
            self.sb[0].copy_(self.fixed_bins[0][0]), self.eb[0].copy_(self.fixed_bins[0][1])
        for lvl in (0, 1):
            # density of the level's samples, then get_weights + PDF resampling in one launch
            self._density_net_fwd(lvl, self.props[lvl], weights=False)
            call("b2n_weights_pdf_sample", ptr(self.sb[lvl]), ptr(self.eb[lvl]), ptr(self.dens[lvl]), ptr(self.u_base[lvl]),
                 NULL if ev else ptr(self.jitter[lvl + 1]), 0,
                 ptr(self.nears), ptr(self.fars), R, self.S[lvl], self.S[lvl + 1] + 1, 1.0, _off(self.hyper, 3), 0.01, 1e-5,
                 lib.SPACING[self.spacing], ptr(self.w[lvl]), ptr(self.sb[lvl + 1]), ptr(self.eb[lvl + 1]), st())
            if self.fixed_bins is not None:  # staged parity tests: every level sees the reference's recorded samples
                self.sb[lvl + 1].copy_(self.fixed_bins[lvl + 1][0]), self.eb[lvl + 1].copy_(self.fixed_bins[lvl + 1][1])
            if ev:
                eb = self.eb[lvl]
                call("b2n_composite_fwd", NULL, ptr(self.w[lvl]), ptr(eb), _off(eb, 1), self.S[lvl] + 1, R, self.S[lvl],
                     lib.BG_NONE, NULL, 0, NULL, NULL, NULL, ptr(self.prop_depth[lvl]), NULL, st())

    def _fork_head_static(self) -> None:
        if self._head_static_join is not None or not self._fork or self.eval_mode or self.hin_stride % 4 != 0:
            return

        def head_static() -> None:
            call("b2n_sh_fwd", ptr(self.directions), self.R, 4, 1, ptr(self.sh), stream())
            self._head_input(1)

        self._head_static_join = self._forked(0, head_static)

    def _head_input(self, part: int) -> None:
        """Colour-head input rows [SH | geo features | appearance embedding] (nerfacto_field.py:234-310); part 1 = the
        columns that are constant along a ray, 2 = the geo features, 0 = all."""
        R, S2, ev = self.R, self.S[2], self.eval_mode
        if self.emb is None:
            emb_ptr, emb_mode = NULL, 0
        elif ev:  # nerfacto_field.py:250-261: mean embedding (pre-averaged into emb_mean by the caller) or zeros
            emb_ptr, emb_mode = (ptr(self.emb_mean), 2) if self.model.field.use_average_appearance_embedding else (NULL, 0)
        else:
            emb_ptr, emb_mode = ptr(self.emb), 1
        call("b2n_head_input_fwd_part", ptr(self.sh), self.n_sh, ptr(self.h[2]), self.h[2].shape[1], self.geo, emb_ptr,
             ptr(self.cams, torch.int64), self.n_emb, emb_mode, R, S2, ptr(self.hin), self.hin_stride, part, stream())

    def _forward_main(self) -> None:
        """Main field on the final samples: positions, hash grid, base MLP, colour head (+ weights and renderers unless the
        fused ray tail does them)."""
        R, S0, S1, S2 = self.R, *self.S
        st = stream
        ev = self.eval_mode
        if self._prologue_join is not None:
            torch.cuda.current_stream().wait_event(self._prologue_join)
            self._prologue_join = None
        if self.tma_weights and self._pack_late:
            self._pack_weights()
        self._fork_head_static()  # (no-op when the branch is already running)
        N2 = R * S2
        eb2 = self.eb[2]
        box = lib.host_floats(self.aabb)
        call("b2n_positions_fwd", ptr(self.origins), ptr(self.directions), ptr(eb2), _off(eb2, 1), S2 + 1, R, S2,
             int(self.contraction), C.cast(box, C.c_void_p), ptr(self.x[2]), ptr(self.sel[2], torch.uint8), st())
        call("b2n_hashgrid_fwd", C.byref(self.base.grid.c), ptr(self.x[2]), ptr(self.base.table), N2, ptr(self.enc[2]), NULL, st())
        mb, gb = self.base.structs()
        self._mlp_fwd(mb, self.enc[2], self.enc[2].shape[1], N2, self.h[2], self.hid[2], self.base.spec)
        bw = self.h[2].shape[1]
        tail = self.fused_tail and not ev  # density activation, weights and renderers run inside b2n_nerfacto_ray_tail
        if not tail:
            call("b2n_density_act_fwd", ptr(self.h[2]), bw, ptr(self.sel[2], torch.uint8), N2, self.avg, ptr(self.dens[2]), st())
        if self._head_static_join is not None:  # SH + embedding columns of the head input were written by a forked branch
            torch.cuda.current_stream().wait_event(self._head_static_join)
            self._head_static_join = None
            self._head_input(2)
        else:
            call("b2n_sh_fwd", ptr(self.directions), R, 4, 1, ptr(self.sh), st())
            self._head_input(0)
        mh = self.head_spec.struct(self.head_w, self.head_b)
        gh = B2nMlpGrad()
        for i, (w, b) in enumerate(zip(self.head_w, self.head_b)):
            gh.dw[i], gh.db[i] = ptr(w.grad).value, ptr(b.grad).value
        self._mlp_fwd(mh, self.hin, self.hin_stride, N2, self.rgb, self.hid_head, self.head_spec)
        bg_mode, bg_ptr, _keep = F._bg_args(self.bg)
        if not tail:
            call("b2n_weights_fwd", ptr(eb2), _off(eb2, 1), S2 + 1, ptr(self.dens[2]), R, S2, ptr(self.w[2]), st())
            call("b2n_composite_fwd", ptr(self.rgb), ptr(self.w[2]), ptr(eb2), _off(eb2, 1), S2 + 1, R, S2, bg_mode, bg_ptr,
                 1 if ev else 0, ptr(self.rgb_out), ptr(self.acc), ptr(self.depth_exp), ptr(self.depth_med), NULL, st())
        self._fw = (mb, gb, mh, gh, bw, bg_mode, bg_ptr, _keep)

    def _body(self, update_props: bool) -> None:
        self._body_a(update_props)
        self._body_b(update_props)

    def _body_a(self, update_props: bool) -> None:
        """Step prologue (draws, zeroing) and the proposal sampling: touches no field parameter."""
        R = self.R
        st = stream
        g = self.optim.flat_grad
        late = self._pack_late

        def prologue() -> None:  # nothing before the main-field forward needs the packed MLP weights or the gradient buffer
            call("b2n_zero_async", ptr(g), g.numel() * g.element_size(), stream())
            if self.tma_weights and not late:
                self._pack_weights()

        if self._fork:
            self._prologue_join = self._forked(1, prologue)
        else:
            call("b2n_zero_async", ptr(g), g.numel() * g.element_size(), st())
        draw = self.fixed_jitter is None
        call("b2n_step_begin", ptr(self.jitter_all) if draw else NULL, 3 * R if draw else 0, ptr(self.rng_state, torch.int64),
             ptr(self.losses), self.losses.numel(), ptr(self.d_rays) if self.camopt is not None else NULL,
             self.d_rays.numel() if self.camopt is not None else 0, st())
        if not draw:  # tests: replay recorded stratified draws
            for j, src in zip(self.jitter, self.fixed_jitter):
                j.copy_(src)
        self._forward_props()
        if late and self._prologue_join is not None:  # this piece is captured on its own: join its branch here
            torch.cuda.current_stream().wait_event(self._prologue_join)
            self._prologue_join = None

    def _body_b(self, update_props: bool) -> None:
        """Main-field forward, the per-ray middle, the main backward."""
        R, S0, S1, S2 = self.R, *self.S
        cfg = self.cfg
        st = stream
        self._forward_main()
        mb, gb, mh, gh, bw, bg_mode, bg_ptr, _keep = self._fw
        N2 = R * S2
        eb2 = self.eb[2]
        il = cfg.interlevel_loss_mult / float(R * S2)
        dm = cfg.distortion_loss_mult / float(R)
        if self.fused_tail:
            # ---------------- weights, renderers, losses and their backward down to the density pre-activation: one launch
            arr = lambda ts: (C.c_void_p * len(ts))(*[ptr(t).value for t in ts])
            self._tail_keep = (arr(self.sb[:2]), arr(self.eb[:2]), arr(self.w[:2]), arr(self.dens[:2]), arr(self.d_w[:2]),
                               arr(self.d_dens[:2]), arr(self.rows))
            a_sb, a_eb, a_w, a_dens, a_dw, a_ddens, a_rows = self._tail_keep
            call("b2n_nerfacto_ray_tail", R, S2, S0, S1, ptr(self.sb[2]), ptr(eb2), ptr(self.h[2]), bw, ptr(self.sel[2], torch.uint8),
                 self.avg, ptr(self.rgb), ptr(self.gt), bg_mode, bg_ptr, 1.0, il, dm, a_sb, a_eb, a_w, a_dens,
                 a_dw if update_props else NULL, a_ddens if update_props else NULL, ptr(self.dens[2]), ptr(self.w[2]),
                 ptr(self.rgb_out), ptr(self.acc), ptr(self.depth_exp), ptr(self.depth_med), ptr(self.d_rgb), ptr(self.d_w[2]),
                 ptr(self.d_w_dist), ptr(self.d_hpre), a_rows, st())
        else:
            self._tail_unfused(update_props, il, dm)
        self._joins = []
        if self._fork_props and update_props and self.fused_tail:
            # proposal backward: its inputs (d_density of both levels) are complete -> parallel branch
            def prop_branch() -> None:
                for lvl in (0, 1):
                    self._density_net_bwd(lvl, self.props[lvl], None)
                self._adam(update_props, only_props=True)  # their gradients are complete: step them here, off the main chain
                self._props_stepped = True

            self._joins.append(self._forked(0, prop_branch))
        self._mlp_bwd(mh, gh, self.hin, self.hin_stride, self.rgb, self.hid_head, self.d_rgb, N2, self.d_hin, self.hin_stride,
                      self.head_spec)
        emb_grad = ptr(self.emb.grad) if self.emb is not None else NULL

        def head_in_bwd(part: int) -> None:
            call("b2n_head_input_bwd_part", ptr(self.d_hin), self.hin_stride, self.n_sh, self.geo, self.n_emb, ptr(self.d_hpre),
                 ptr(self.cams, torch.int64), R, S2, ptr(self.d_h[2]), bw, emb_grad, part, stream())

        emb_join = None
        if self._fork and self.emb is not None:
            head_in_bwd(1)                                        # d(base output): the base MLP's backward waits for it
            emb_join = self._forked(1, lambda: head_in_bwd(2))    # embedding-row gradients: beside the base backward
        else:
            head_in_bwd(0)
        self._mlp_bwd(mb, gb, self.enc[2], self.enc[2].shape[1], self.h[2], self.hid[2], self.d_h[2], N2, self.d_enc[2],
                      self.d_enc[2].shape[1], self.base.spec)
        def main_dx() -> None:
            # main level: d enc -> d x (gather-only kernel) -> contraction Jacobian -> the ray's d origin / d direction
            box = lib.host_floats(self.aabb)
            call("b2n_hashgrid_dx", C.byref(self.base.grid.c), ptr(self.x[2]), ptr(self.base.table), ptr(self.d_enc[2]), N2,
                 ptr(self.d_x2), stream())
            call("b2n_positions_bwd", ptr(self.origins), ptr(self.directions), ptr(eb2), _off(eb2, 1), S2 + 1, R, S2,
                 int(self.contraction), C.cast(box, C.c_void_p), ptr(self.d_x2), 2 if self._fork else 1, ptr(self.d_rays[0]),
                 ptr(self.d_rays[1]), stream())

        dx_join = None
        if self.camopt is not None and self._fork:
            dx_join = self._forked(1, main_dx)  # gather-bound, next to the atomics-bound scatter below
        call("b2n_hashgrid_bwd", C.byref(self.base.grid.c), ptr(self.x[2]), ptr(self.base.table), ptr(self.d_enc[2]), N2,
             ptr(self.base.table.grad), NULL, st())
        if self.camopt is not None:
            if not self._fork:
                main_dx()
            cc = self.camopt.config  # regulariser (camera_optimizers.py:155-162): value into losses[4], gradient into the poses
            call("b2n_pose_regularizer", ptr(self.cam_pose), self.cam_pose.shape[0], float(cc.trans_l2_penalty),
                 float(cc.rot_l2_penalty), 1.0, _off(self.losses, 4), ptr(self.cam_pose.grad), st())
        if self.fused_tail:
            call("b2n_loss_finalize", self._tail_keep[6], R, il, dm, 1.0 / float(3 * R), ptr(self.losses), st())
        else:
            call("b2n_loss_total", ptr(self.losses), 3, _off(self.losses, 4), _off(self.losses, 3), st())
        for ev_ in (dx_join, emb_join):
            if ev_ is not None:
                torch.cuda.current_stream().wait_event(ev_)

    def _tail_unfused(self, update_props: bool, il: float, dm: float) -> None:
        """The per-ray middle of the step, one operator per launch (what b2n_nerfacto_ray_tail fuses)."""
        R, S0, S1, S2 = self.R, *self.S
        st = stream
        mb, gb, mh, gh, bw, bg_mode, bg_ptr, _keep = self._fw
        N2 = R * S2
        eb2 = self.eb[2]
        # ---------------- losses (+ their gradients)
        call("b2n_mse_fwd_bwd", ptr(self.rgb_out), ptr(self.gt), 3 * R, 1.0, ptr(self.losses), ptr(self.d_rgb_out), st())
        for lvl in (0, 1):
            call("b2n_interlevel_fwd_bwd", ptr(self.sb[2]), ptr(self.w[2]), ptr(self.sb[lvl]), ptr(self.w[lvl]), R, S2,
                 self.S[lvl], il, ptr(self.rows[lvl]), ptr(self.d_w[lvl]) if update_props else NULL, st())
            call("b2n_sum_rows", ptr(self.rows[lvl]), R, il, _off(self.losses, 1), st())
        call("b2n_distortion_fwd_bwd", ptr(self.sb[2]), ptr(self.w[2]), R, S2, dm, ptr(self.rows[2]), ptr(self.d_w_dist), st())
        call("b2n_sum_rows", ptr(self.rows[2]), R, dm, _off(self.losses, 2), st())
        # ---------------- backward: main field
        call("b2n_composite_bwd", ptr(self.rgb), ptr(self.w[2]), ptr(eb2), _off(eb2, 1), S2 + 1, ptr(self.d_rgb_out), NULL, NULL,
             R, S2, bg_mode, bg_ptr, ptr(self.d_rgb), ptr(self.d_w[2]), st())
        call("b2n_add_inplace", ptr(self.d_w[2]), ptr(self.d_w_dist), self.d_w[2].numel(), st())
        call("b2n_weights_bwd", ptr(eb2), _off(eb2, 1), S2 + 1, ptr(self.dens[2]), ptr(self.d_w[2]), R, S2, ptr(self.d_dens[2]), st())
        call("b2n_density_act_bwd", ptr(self.h[2]), bw, ptr(self.sel[2], torch.uint8), ptr(self.d_dens[2]), N2, self.avg,
             ptr(self.d_hpre), 1, st())

    def _forked(self, k: int, fn) -> torch.cuda.Event:
        """Run fn() on side stream k after everything issued so far on the current stream; returns the event to join on."""
        cur, side = torch.cuda.current_stream(), self._side[k]
        start = torch.cuda.Event()
        start.record(cur)
        side.wait_event(start)
        with torch.cuda.stream(side):
            fn()
            done = torch.cuda.Event()
            done.record(side)
        return done

    def _body_props(self, update_props: bool) -> None:
        """backward of the proposal networks (only the interlevel loss reaches them)."""
        forked_props = self._fork_props and update_props and self.fused_tail
        for ev in self._joins:  # branches forked in _body
            torch.cuda.current_stream().wait_event(ev)
        self._joins = []
        if update_props and not forked_props:
            for lvl in (0, 1):
                self._density_net_bwd(lvl, self.props[lvl], None)
        if self.camopt is not None:  # all three levels have added their d origin / d direction: on to the poses
            call("b2n_pose_apply_bwd", ptr(self.cam_pose), ptr(self.cams, torch.int64), ptr(self.cam_frozen, torch.uint8),
                 ptr(self.directions_in), ptr(self.d_rays[0]), ptr(self.d_rays[1]), self.R, ptr(self.cam_pose.grad), stream())

    def _adam(self, update: bool = True, only_props: bool = False) -> None:
        """Fused Adam over the flat buffer.  Frozen proposal networks are NOT stepped (reference: their grads are None
        and engine/optimizers.py:155 skips the group), and when they are, they use their own bias-correction count.
        `only_props`: just the proposal group (stepped at the end of the forked proposal branch, beside the main backward);
        the regular call then leaves that group out."""
        o = self.optim
        runs = []
        props_done = False
        if not only_props:
            props_done, self._props_stepped = self._props_stepped, False
        for name, a, b in o.segments:
            is_prop = name == "proposal_networks"
            if is_prop and (not update or props_done):
                continue
            if only_props and not is_prop:
                continue
            slot = 8 if name == "camera_opt" else (4 if (is_prop and not self.always_update) else 0)
            if runs and runs[-1][1] == a and runs[-1][2] == slot:
                runs[-1][1] = b
            else:
                runs.append([a, b, slot])
        if self._shard is not None:
            # sharded update: of the field segment [0, grad_split) this rank steps its own slice (whose gradient sums the
            # reduce-scatter left here) and the short tail that went through the all-reduce; the all-gather that follows
            # brings the other ranks' slices of the updated parameters
            chunk, m0, m1, tail = self._shard
            split_runs = []
            for a, b, slot in runs:
                if b <= self.grad_split:
                    for lo, hi in ((max(a, m0), min(b, m1)), (max(a, tail), b)):
                        if hi > lo:
                            split_runs.append([lo, hi, slot])
                else:
                    assert a >= self.grad_split
                    split_runs.append([a, b, slot])
            runs = split_runs
        for a, b, slot in runs:
            call("b2n_adam_step_dev", _off(o.flat, a), _off(o.flat_grad, a), _off(o.exp_avg, a), _off(o.exp_avg_sq, a),
                 b - a, _off(self.hyper, slot), float(o.betas[0]), float(o.betas[1]), float(o.eps), stream())

    # ------------------------------------------------------------------------------------------------
    def set_batch(self, origins: Tensor, directions: Tensor, camera_indices: Tensor, gt_rgb: Tensor) -> None:
        """Copy one ray batch into the static input buffers (H2D when the sources are pinned host tensors)."""
        self.origins_in.copy_(origins, non_blocking=True)
        self.directions_in.copy_(directions, non_blocking=True)
        self.cams.copy_(camera_indices.reshape(-1), non_blocking=True)
        self.gt.copy_(gt_rgb, non_blocking=True)

    def pack_batch(self, origins: Tensor, directions: Tensor, camera_indices: Tensor, gt_rgb: Tensor) -> Tensor:
        """A pinned host blob in the layout of the static input buffer (what a data loader would hand over)."""
        R = self.R
        blob = torch.empty(11 * R, dtype=torch.float32).pin_memory()
        blob[: 2 * R].view(torch.int64).copy_(camera_indices.reshape(-1))
        blob[2 * R: 5 * R].view(R, 3).copy_(origins)
        blob[5 * R: 8 * R].view(R, 3).copy_(directions)
        blob[8 * R: 11 * R].view(R, 3).copy_(gt_rgb)
        return blob

    def set_batch_packed(self, blob: Tensor) -> None:
        """One H2D (or D2D) copy of a batch laid out by pack_batch."""
        self.inputs.copy_(blob, non_blocking=True)

    def _anneal(self, step: int) -> float:
        c = self.cfg
        if not c.use_proposal_weight_anneal:
            return 1.0
        frac = min(max(step / c.proposal_weights_anneal_max_num_iters, 0.0), 1.0)
        b = c.proposal_weights_anneal_slope
        return b * frac / ((b - 1) * frac + 1)

    def _update_due(self, step: int) -> bool:
        if self.always_update:
            return True
        return self._steps_since_update > self.model.proposal_sampler.update_sched(step) or step < 10

    def step(self) -> Tensor:
        """One optimisation step on the batch currently in the static buffers.  Returns the loss vector
        [rgb, interlevel, distortion, total] (device tensor, valid after the stream is synchronised)."""
        t = self.step_count
        o = self.optim
        lr = o.lr_schedule(t) if o.lr_schedule is not None else o.lr
        world = getattr(self.allreduce, "world", 1) if self.allreduce is not None else 1
        slot = t % self.HYPER_SLOTS
        if self._hyper_events[slot] is not None:
            self._hyper_events[slot].synchronize()
        pt = self._prop_steps
        clr = self.camera_lr_schedule(t) if self.camera_lr_schedule is not None else self.camera_lr
        self._hyper_np[slot] = (lr / (1.0 - o.betas[0] ** (t + 1)), 1.0 / math.sqrt(1.0 - o.betas[1] ** (t + 1)),
                                1.0 / world, self._anneal(t),
                                lr / (1.0 - o.betas[0] ** (pt + 1)), 1.0 / math.sqrt(1.0 - o.betas[1] ** (pt + 1)),
                                1.0 / world, 0.0,
                                clr / (1.0 - o.betas[0] ** (t + 1)), 1.0 / math.sqrt(1.0 - o.betas[1] ** (t + 1)),
                                1.0 / world, 0.0)
        self.hyper.copy_(self._hyper_ring[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_events[slot] = ev
        update = self._update_due(t)
        overlap = self.allreduce is not None and world > 1
        self._fork, self._fork_props = self.concurrent, self.concurrent and not overlap
        shard = (overlap and self.sharded_update and self.grad_split is not None and self.fused_tail
                 and hasattr(self.allreduce, "start_reduce_scatter") and self.allreduce.shard_chunk(self.grad_split) > 0)
        if shard:
            return self._step_sharded(update, world)
        self.flush()
        if self.use_graph and update not in self._graphs:
            self._capture(update, split=overlap)
        if self.use_graph and not overlap:
            # single process: the whole step is ONE graph launch
            self._graphs[update][3].replay()
            return self._finish_step(update)

        def run(i: int) -> None:
            if self.use_graph:
                g = self._graphs[update][i]
                if g is not None:
                    g.replay()
            else:
                (self._body, self._body_props, self._adam)[i](update)

        # [forward + losses + main-field backward] -> start summing the field's gradient segment over NVLink while
        # the [proposal backward] runs -> sum the proposal segment -> [Adam].  One process per GPU, two collectives.
        run(0)
        split = self.grad_split if overlap else None
        h_field = self.allreduce.start(o.flat_grad[:split]) if split is not None else None
        run(1)
        h_rest = None
        if overlap and split is None:
            h_rest = self.allreduce.start(o.flat_grad)  # unknown layout: one collective after the whole backward
        elif overlap and update:
            h_rest = self.allreduce.start(o.flat_grad[split:])
        elif overlap and self._cam_seg is not None:  # frozen proposals: nothing to sum there; the poses still train
            h_rest = self.allreduce.start(o.flat_grad[self._cam_seg[0]: self._cam_seg[1]])
        if overlap:
            self.allreduce.finish(h_field, h_rest)
        run(2)
        return self._finish_step(update)

    def flush(self) -> None:
        """Make the current stream wait for a pending parameter all-gather (sharded update): call before anything other
        than step() reads the field parameters (evaluation, checkpoints)."""
        if self._h_ag is not None:
            self._h_ag.wait()
            self._h_ag = None

    def _step_sharded(self, update: bool, world: int) -> Tensor:
        """One process per GPU, the field segment (the 67 MB hash table) updated in shards:
            [draws + proposal sampling]            | all-gather of the field parameters of the PREVIOUS step still landing
            wait all-gather -> [main forward + ray tail + main backward]
            reduce-scatter(field gradients)        | [proposal + pose backward]
            all-reduce(field tail + camera + proposal gradients)
            [Adam: own field slice, field tail, camera, proposals] -> all-gather(field parameters), asynchronous.
        Same wire bytes as one all-reduce of the buffer, but only the reduce-scatter half sits between backward and Adam,
        the other half overlaps the next step's proposal sampling, and Adam touches 1/world of the table."""
        o, ar = self.optim, self.allreduce
        rank = torch.distributed.get_rank(ar.group)
        chunk = ar.shard_chunk(self.grad_split)
        self._shard = (chunk, rank * chunk, (rank + 1) * chunk, world * chunk)
        key = ("shard", update)
        if self.use_graph and key not in self._graphs:
            self._capture_sharded(update)

        def run(i: int) -> None:
            if self.use_graph:
                g = self._graphs[key][i]
                if g is not None:
                    g.replay()
            else:
                self._pack_late = True
                (self._body_a, self._body_b, self._body_props, self._adam)[i](update)
                self._pack_late = False

        run(0)
        self.flush()  # the field parameters of the previous step are complete before the main forward reads them
        run(1)
        h_rs = ar.start_reduce_scatter(o.flat_grad[: world * chunk])
        run(2)
        end = o.flat_grad.numel() if update else (self._cam_seg[1] if self._cam_seg is not None and
                                                   self._cam_seg[0] == self.grad_split else self.grad_split)
        h_rest = ar.start(o.flat_grad[world * chunk: end]) if end > world * chunk else None
        ar.finish(h_rs, h_rest)
        run(3)
        self._h_ag = ar.start_all_gather(o.flat[: world * chunk])
        self._shard = None
        return self._finish_step(update)

    def _capture_sharded(self, update: bool) -> None:
        """Four graphs: [prologue + proposal sampling] [main forward .. main backward] [proposal + pose backward] [Adam]."""
        self._fork, self._fork_props, self._pack_late = self.concurrent, False, True
        saved = [t.clone() for t in (self.optim.flat, self.optim.exp_avg, self.optim.exp_avg_sq)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body_a(update), self._body_b(update), self._body_props(update), self._adam(update)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for dst, src in zip((self.optim.flat, self.optim.exp_avg, self.optim.exp_avg_sq), saved):
            dst.copy_(src)  # the warm-up pass must not count as an optimisation step
        g_a, g_b, g_props, g_adam = (torch.cuda.CUDAGraph() for _ in range(4))
        with torch.cuda.graph(g_a):
            self._body_a(update)
        with torch.cuda.graph(g_b, pool=g_a.pool()):
            self._body_b(update)
        if update or self.camopt is not None:
            with torch.cuda.graph(g_props, pool=g_a.pool()):
                self._body_props(update)
        else:
            g_props = None
        with torch.cuda.graph(g_adam, pool=g_a.pool()):
            self._adam(update)
        self._pack_late = False
        self._graphs[("shard", update)] = (g_a, g_b, g_props, g_adam)

    def _finish_step(self, update: bool) -> Tensor:
        if update:
            self._steps_since_update = 0
            self._prop_steps += 1
        self._steps_since_update += 1
        self.step_count += 1
        self.optim.steps += 1
        for name in self.optim.group_steps:
            if update or name != "proposal_networks":
                self.optim.group_steps[name] += 1
        return self.losses

    def _capture(self, update: bool, split: bool) -> None:
        """Warm the kernels up on a side stream (loads modules, sizes smem attributes), then capture: three graphs
        (forward + main backward | proposal backward | Adam) when collectives run between them, else one."""
        self._fork, self._fork_props = self.concurrent, self.concurrent and not split
        saved = [t.clone() for t in (self.optim.flat, self.optim.exp_avg, self.optim.exp_avg_sq)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body(update)
            self._body_props(update)
            self._adam(update)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for dst, src in zip((self.optim.flat, self.optim.exp_avg, self.optim.exp_avg_sq), saved):
            dst.copy_(src)  # the warm-up pass must not count as an optimisation step
        if not split:
            # single process: no collective between the pieces, so the whole step is ONE graph
            g_all = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_all):
                self._body(update)
                if update or self.camopt is not None:
                    self._body_props(update)
                self._adam(update)
            self._graphs[update] = (None, None, None, g_all)
            return
        g_main, g_props, g_adam = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_main):
            self._body(update)
        if update or self.camopt is not None:
            with torch.cuda.graph(g_props, pool=g_main.pool()):
                self._body_props(update)
        else:
            g_props = None  # nothing to replay when the proposal networks are frozen this step
        with torch.cuda.graph(g_adam, pool=g_main.pool()):
            self._adam(update)
        self._graphs[update] = (g_main, g_props, g_adam, None)
