"""Drop-in replacement for the reference's src/trainer.py `Trainer` (same constructor, `.train()`,
`.iteration`, `.epoch`, snapshot files), with the per-iteration body of trainer.py:271-363 executed as
one fused schedule over the libdcvgan_b200 kernels instead of PyTorch autograd:

  * every forward pass and every RNG draw happens in the reference's order (real idis/vdis/gdis, ggen,
    cgen, fake idis/vdis/gdis; then the G-phase), so seeds and BatchNorm running statistics evolve
    identically - including the D-phase fakes being generated in whatever train/eval mode the generators
    were left in (trainer.py:126-127 quirk), D BatchNorm always in training mode, the swapped update
    gates (:318,:355) and opt_ggen stepping twice (:357,:359);
  * the provably dead work is skipped: no generator backward in the D-phase, no
    discriminator weight gradients in the G-phase;
  * parameters, gradients and Adam moments of each network live in flat fp32 buffers: one Adam launch per
    optimizer step and one NCCL all-reduce per network per phase when torch.distributed is initialised
    (data-parallel replicas, per-replica BatchNorm statistics);
  * losses stay on the device and reach `logger.update()` in order when the log interval fires, instead of
    four host syncs per iteration (trainer.py:326-328,363).
"""
import copy
import os
import shutil
import weakref
from pathlib import Path
from typing import Any, Dict

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from . import _lib, engine, ops
from .generator import current_device
from .ops import Act

import enum


class MetricType(enum.IntEnum):
    """Same members and values as the reference's logger.MetricType (logger.py:15-23, an IntEnum): the reference's Logger
    compares metric types by integer value, so this works whether or not its `logger` module is importable."""
    Integer = 1
    Float = 2
    Loss = 3
    Time = 4


class _FlatNet:
    """Flat fp32 storage for one network's parameters, gradients and Adam moments."""

    def __init__(self, model, opt, grad_storage=None):
        self.model, self.opt = model, opt
        self.params = [p for p in model.parameters()]
        n = sum(p.numel() for p in self.params)
        # 4-element alignment per tensor keeps every view 16-byte aligned for the vectorised Adam path
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        dev = self.params[0].device
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        # grad_storage: a slice of a buffer shared by the networks of one phase, so that the phase needs ONE all-reduce
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev) if grad_storage is None else grad_storage
        assert self.flat_g.numel() == total
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_of = {}
        self.numel = n
        self.offs = offs
        for p, o in zip(self.params, offs):
            k = p.numel()
            self.flat_p[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + k].view(p.shape)
            g = self.flat_g[o:o + k].view(p.shape)
            p.grad = g
            self.grad_of[p] = g
            st = opt.state[p]
            if "exp_avg" in st:  # resume from an optimizer that already stepped
                self.flat_m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                self.flat_v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            st["exp_avg"] = self.flat_m[o:o + k].view(p.shape)
            st["exp_avg_sq"] = self.flat_v[o:o + k].view(p.shape)
            if "step" not in st:
                st["step"] = torch.tensor(0.0)
        # {int64 step; float lr/bias_correction1; float 1/sqrt(bias_correction2)} on the parameters' device: the step
        # counter never passes through the host, so the fused step can be replayed as a CUDA graph
        step0 = int(float(opt.state[self.params[0]]["step"]))
        self.step_state = torch.zeros(2, dtype=torch.int64, device=dev)
        self.step_state[0] = step0

    @staticmethod
    def padded_numel(model):
        return sum((p.numel() + 3) // 4 * 4 for p in model.parameters())

    def still_bound(self):
        """False when a parameter no longer aliases its slice of flat_p (model.to('cpu').to(device), load_state_dict with
        assign=True, ...): Adam and captured graphs would then update memory the forward passes no longer read."""
        el = self.flat_p.element_size()
        return all(p.data_ptr() == self.flat_p.data_ptr() + o * el and p.grad is self.grad_of.get(p)
                   for p, o in zip(self.params, self.offs))

    def rebind(self):
        """adopt the parameters' current values and make them views of the flat buffers again"""
        for p, o in zip(self.params, self.offs):
            k = p.numel()
            view = self.flat_p[o:o + k].view(p.shape)
            if p.data_ptr() != view.data_ptr():
                view.copy_(p.data.to(view.device))
                p.data = view
            p.grad = self.grad_of[p]

    def adam_step(self, grad_scale=1.0):
        """torch.optim.Adam.step() on the flat buffers (train.py:170-176 hyper-parameters)."""
        g = self.opt.param_groups[0]
        b1, b2 = g["betas"]
        _lib.check(_lib.lib().dcv_adam_flat_dev(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                                self.flat_v.data_ptr(), self.flat_p.numel(), g["lr"], b1, b2, g["eps"],
                                                g["weight_decay"], self.step_state.data_ptr(), grad_scale,
                                                torch.cuda.current_stream().cuda_stream))
        engine.invalidate_packed(self.params)   # the packed bf16 copies of these weights are now stale

    def sync_optimizer_state(self):
        """publish the device-side step count into opt.state[p]['step'] (host sync; call before saving / inspecting)"""
        step = float(int(self.step_state[0].item()))
        for p in self.params:
            self.opt.state[p]["step"] = torch.tensor(step)


def _cpu_state(osd):
    """optimizer.state_dict() with every tensor detached, cloned and moved to the CPU"""
    def conv(v):
        if isinstance(v, torch.Tensor):
            return v.detach().clone().cpu()
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return type(v)(conv(x) for x in v)
        return v
    return conv(osd)


def dp_world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def dp_sync_params(flat_nets):
    """Identical replicas at start: rank 0's flat parameter buffers win."""
    if dp_world() > 1:
        for f in flat_nets:
            dist.broadcast(f.flat_p, 0)


def dp_allreduce_grads(flat_nets):
    """One sum all-reduce per network over its flat gradient buffer (NCCL over NVLink on the GPU box; gloo in the
    CPU tests).  The 1/world average is folded into the Adam kernel's grad_scale.  Returns that scale."""
    world = dp_world()
    if world > 1:
        for f in flat_nets:
            dist.all_reduce(f.flat_g, op=dist.ReduceOp.SUM)
    return 1.0 / world


class GradReducer:
    """Bucketed gradient all-reduce overlapped with the rest of the backward pass.

    One bucket = one network's flat fp32 gradient buffer.  `launch(flat)` is called right after that network's last
    weight gradient: the all-reduce is enqueued on a communication stream that first waits for an event recorded on
    the compute stream at that point, so NCCL runs over NVLink while the compute stream continues with the next
    network's backward (D phase: idis -> vdis -> gdis; G phase: cgen's reduction overlaps ggen's backward).  `wait()`
    makes the compute stream wait for every outstanding bucket before the Adam kernels read the gradients.  Both
    calls are stream/event operations only, so they are captured into the CUDA graph of the step as a fork/join.
    On CPU tensors (gloo tests) the reduction is synchronous."""

    def __init__(self):
        self.stream = None
        self.pending = []

    def launch(self, flat):
        if dp_world() == 1:
            return
        g = flat.flat_g
        if not g.is_cuda:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.pending.append(done)

    def wait(self):
        cur = torch.cuda.current_stream() if self.pending else None
        for ev in self.pending:
            cur.wait_event(ev)
        self.pending = []
        return 1.0 / dp_world()


def dp_seed(seed):
    """Replica r draws its latents / Noise / Dropout from seed + r (each replica is a valid reference step on its
    own shard); returns the seed used."""
    r = dist.get_rank() if dp_world() > 1 else 0
    torch.manual_seed(seed + r)
    np.random.seed(seed + r)
    return seed + r


class Trainer(object):
    def __init__(self, dataloader, logger, models: Dict[str, nn.Module], optimizers: Dict[str, Any], loss,
                 configs: Dict[str, Any]):
        self.dataloader = dataloader
        self.logger = logger
        self.models = models
        self.optimizers = optimizers
        self.loss = loss
        self.configs = configs
        self.device = current_device()
        self.geometric_info = configs["geometric_info"]["name"]
        self.num_log, self.rows_log, self.cols_log = 25, 5, 5
        self.eval_batchsize = configs["evaluation"]["batchsize"]
        self.eval_num_samples = configs["evaluation"]["num_samples"]
        self.eval_metrics = configs["evaluation"]["metrics"]
        self.use_gdis = "gdis" in models and models["gdis"] is not None and configs.get("gdis", {}).get("enabled", True)
        if not hasattr(loss, "KIND_REAL"):
            raise TypeError("dcvgan_b200.trainer.Trainer needs a dcvgan_b200.loss.AdversarialLoss / HingeLoss instance")

        self.model_snapshots_path = Path(self.logger.path) / "models"
        self.model_snapshots_path.mkdir(parents=True, exist_ok=True)
        if configs.get("config_path") and Path(configs["config_path"]).exists():
            shutil.copy(configs["config_path"], str(Path(self.logger.path) / "config.yml"))

        self.iteration: int = 0
        self.epoch: int = 0
        self._flat = None
        self._wcache = {}
        self.replayed_launches = 0  # dcv kernels executed through graph replays (dcv_launch_count() only sees eager launches)
        self._graphs = {}           # (upd_d, upd_g, ggen.training, cgen.training) -> [eager runs so far, CUDAGraph, losses]
        self.use_cuda_graph = os.environ.get("DCV_NO_GRAPH", "0") != "1"
        self._pending = []          # device-side loss records waiting for the next log flush
        self._zero_pool = ops.ZeroPool() if os.environ.get("DCV_NO_ZERO_POOL", "0") != "1" else None
        self._reducer = GradReducer()
        self._sm_reserved = False
        self._copy_stream = None
        self.on_log_samples = None  # optional hooks for the (out-of-scope) visual logging / IS-FID evaluation
        self.on_evaluate = None
        self.save_classobj()

    # ------------------------------------------------------------------ periodic side work (trainer.py:70-224)
    def save_classobj(self):
        """whole-module pickles, `<log>/models/{name}_model.pth` (trainer.py:70-76)"""
        for name, _model in self.models.items():
            model = copy.deepcopy(_model).cpu()
            torch.save(model, self.model_snapshots_path / f"{name}_model.pth")

    def save_params(self):
        """state_dict snapshots `{name}_params_{iteration:05d}.pth` (trainer.py:78-86)"""
        if dist.is_initialized() and dist.get_rank() != 0:
            return
        for name, model in self.models.items():
            sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
            torch.save(sd, self.model_snapshots_path / f"{name}_params_{self.iteration:05d}.pth")
        # Beyond the reference (which cannot resume): the Adam state of every network and the loop counters.
        self.sync_optimizer_state()
        optim = {name: _cpu_state(opt.state_dict()) for name, opt in self.optimizers.items() if name in self.models}
        torch.save({"iteration": self.iteration, "epoch": self.epoch, "optimizers": optim},
                   self.model_snapshots_path / f"optim_{self.iteration:05d}.pth")

    def load_snapshot(self, iteration, path=None):
        """Resume from the snapshot `save_params()` wrote at `iteration`: model parameters / BatchNorm buffers
        ({name}_params_{iteration:05d}.pth, the reference's format, infer.py:14-38) plus optim_{iteration:05d}.pth."""
        path = Path(path) if path is not None else self.model_snapshots_path
        for name, model in self.models.items():
            model.load_state_dict(torch.load(path / f"{name}_params_{iteration:05d}.pth", map_location="cpu"))
        st = torch.load(path / f"optim_{iteration:05d}.pth", map_location="cpu")
        for name, osd in st["optimizers"].items():
            self.optimizers[name].load_state_dict(osd)
        self.iteration, self.epoch = int(st["iteration"]), int(st["epoch"])
        self._flat = None            # re-adopt parameters and optimizer state into fresh flat buffers at the next step
        self._graphs.clear()
        self._wcache.clear()

    def log_hparams(self):
        def flat(item, key):
            if type(item) != dict:
                return {key: str(item)}
            out = {}
            for k, v in item.items():
                out.update(flat(v, k if key == "" else key + "/" + k))
            return out
        fn = getattr(self.logger, "tf_log_hparams", None)
        if fn is not None:
            fn(flat(self.configs, ""))

    def log_samples(self, ggen, cgen, iteration):
        """Visual logging is outside the hot path; what matters to the step is that the generators are left in
        eval mode afterwards (trainer.py:126-127), which the next D-phase then observes."""
        ggen.eval()
        cgen.eval()
        if self.on_log_samples is not None:
            self.on_log_samples(self, ggen, cgen, iteration)

    def evaluate(self, ggen, cgen):
        if self.on_evaluate is not None:
            ggen.eval()
            cgen.eval()
            self.on_evaluate(self, ggen, cgen)

    # ------------------------------------------------------------------ fused step
    def _prepare(self):
        if self._flat is not None:
            return
        _lib.require_device()
        names = ["ggen", "cgen", "idis", "vdis"] + (["gdis"] if self.use_gdis else [])
        for n in names:
            self.models[n].to(self.device)
        # gradient buffers of the networks that update together share one allocation (D: idis|vdis[|gdis], G: cgen|ggen):
        # data-parallel runs can then reduce a phase with a single collective (DCV_DP_OVERLAP=2)
        self._flat, self._phase_grads = {}, {}
        for phase, members in (("d", names[2:]), ("g", ["cgen", "ggen"])):
            sizes = [_FlatNet.padded_numel(self.models[n]) for n in members]
            buf = torch.zeros(sum(sizes), dtype=torch.float32, device=self.device)
            self._phase_grads[phase] = buf
            o = 0
            for n, k in zip(members, sizes):
                self._flat[n] = _FlatNet(self.models[n], self.optimizers[n], buf[o:o + k])
                o += k
        self._plans = {"ggen": engine.GGenPlan(self.models["ggen"]), "cgen": engine.CGenPlan(self.models["cgen"])}
        for n in names[2:]:
            self._plans[n] = engine.DisPlan(self.models[n], n)
        self._dnames = names[2:]
        self._pack_group = {id(p): n for n in names for p in self.models[n].parameters()}
        self._pack_log = {}
        if os.environ.get("DCV_NO_BATCH_PACK", "0") == "1":
            self._pack_group = None
        self.world = dp_world()
        self.dtype = ops.torch_dtype(self.models["ggen"].precision)
        # production widths (every channel count a multiple of 16): the bf16 path must run on tcgen05 end to end - a layer
        # that would drop to the CUDA-core kernels raises instead (ops.STRICT_TC).  Narrow test widths keep the fallback.
        widths = [self.models["ggen"].ngf, self.models["cgen"].ngf] + [self.models[n].ndf for n in names[2:]]
        self._strict_tc = self.dtype == torch.bfloat16 and all(w % 16 == 0 for w in widths)
        if self.world > 1:
            dp_sync_params(self._flat.values())
            if "seed" in self.configs:
                dp_seed(self.configs["seed"])

    def _check_bound(self):
        """Parameters must still alias the flat buffers (a hook may have moved a model to the CPU and back, as the
        reference's evaluate() does): re-adopt them and drop the captured graphs and packed weights if not."""
        stale = [f for f in self._flat.values() if not f.still_bound()]
        if stale:
            for f in stale:
                f.rebind()
            self._graphs.clear()
            self._wcache.clear()

    def _allreduce(self, names):
        dp_allreduce_grads([self._flat[n] for n in names])

    def _dp_mode(self):
        """DCV_DP_OVERLAP: 2 (default) = ONE in-stream all-reduce per phase over the shared gradient buffer (idis|vdis[|gdis],
        cgen|ggen: two collectives per iteration), 0 = one in-stream all-reduce per network right after its backward,
        1 = side-stream all-reduce per network while the persistent kernels leave DCV_DP_SM_RESERVE SMs free.
        Measured on 8 B200s, mug-depth (profiles/r2l_bench_8gpu_overlap*.json): 10.78 / 10.89 / 10.83 ms per iteration for
        modes 2 / 0 / 1 (1 GPU: 10.37 ms); on 2 GPUs modes 0 / 1: 10.73 / 10.77 ms."""
        return os.environ.get("DCV_DP_OVERLAP", "2")

    def _reduce_launch(self, name):
        """enqueue the all-reduce of one network's gradient bucket behind the kernels issued so far"""
        if self.world <= 1:
            return
        mode = self._dp_mode()
        if mode == "1":
            # side stream + reserved SMs: with one ~220 KB CTA per SM on all 148 SMs NCCL's CTAs only get an SM between two
            # kernels, so the persistent kernels launched until the matching wait use 148 - DCV_DP_SM_RESERVE CTAs
            self._reducer.launch(self._flat[name])
            k = int(os.environ.get("DCV_DP_SM_RESERVE", "8"))
            if k > 0 and not self._sm_reserved:
                _lib.check(_lib.lib().dcv_set_tuning(b"sm_reserve", k))
                self._sm_reserved = True
        elif mode == "0":
            self._allreduce([name])

    def _reduce_wait(self, phase=None):
        if self.world <= 1:
            return
        if self._dp_mode() == "2" and phase is not None:
            dist.all_reduce(self._phase_grads[phase], op=dist.ReduceOp.SUM)
        self._reducer.wait()
        if self._sm_reserved:
            _lib.check(_lib.lib().dcv_set_tuning(b"sm_reserve", 0))
            self._sm_reserved = False

    def _to_clip(self, x, channels=None):
        """Real batch -> channels-last Act (B,T,H,W,C).  Accepted forms:
          * (B,C,T,H,W) float32 - what the reference's dataset yields (dataset.py:128-186);
          * (B,T,H,W,C) uint8   - raw frames as stored on disk: `/ 127.5 - 1` happens on the device (a quarter of the H2D bytes);
          * (B,T,H,W) uint8 / int64 class indices (segmentation): one-hot expansion on the device."""
        if x.dtype == torch.uint8 and x.dim() == 5:
            b, t, h, w, c = x.shape
            a = Act.empty(b, t, h, w, c, self.dtype)
            ops.ingest_u8(x.contiguous(), a)
            return a
        if x.dtype in (torch.uint8, torch.int64) and x.dim() == 4:
            b, t, h, w = x.shape
            a = Act.empty(b, t, h, w, channels, self.dtype)
            ops.ingest_onehot(x.contiguous(), a)
            return a
        b, c, t, h, w = x.shape
        a = Act.empty(b, t, h, w, c, self.dtype)
        ops.to_channels_last(x.float(), a)
        return a

    def _frame(self, clip, t):
        f = Act.empty(clip.n, 1, clip.h, clip.w, clip.c, clip.dtype)
        ops.frame_extract(clip, t, f)
        return f

    def _dis_forward(self, xg, xc, t_rand, save, loss=None):
        """idis on frame t_rand, vdis and gdis on the clips, in the reference's order (trainer.py:299-301).
        loss: None or {name: {'kind', 'out', 'accumulate', 'want_grad'}} - the loss term of each listed discriminator is fused
        into its head launch.  Returns {name: (logits, ctx, head info {'loss_fused', 'dlogits'})}."""
        r = engine.rng()
        loss = loss or {}
        out = {}
        for n, args in (("idis", (self._frame(xg, t_rand), self._frame(xc, t_rand))), ("vdis", (xg, xc)), ("gdis", (xg, None))):
            if n == "gdis" and not self.use_gdis:
                continue
            y, ctx = self._plans[n].forward(args[0], args[1], True, r, save, loss=loss.get(n))
            out[n] = (y, ctx, self._plans[n].last_head)
        return out

    def _generate(self, B, ggen_training, cgen_training, save):
        """ggen.sample_videos + cgen.forward_videos (trainer.py:304-305,344-345) on channels-last buffers."""
        r = engine.rng()
        ggen, cgen = self.models["ggen"], self.models["cgen"]
        T = ggen.video_length
        xg, gctx = self._plans["ggen"].forward(B, ggen_training, self.dtype, r, save)          # (B*T,1,64,64,C)
        z = r.normal((B, cgen.dim_z))                                                           # generator.py:355-359
        zs = z.unsqueeze(1).repeat(1, T, 1).view(B * T, -1)
        xc, cctx = self._plans["cgen"].forward(xg, zs, cgen_training, r, save)                  # (B*T,1,64,64,3)
        return xg, xc, gctx, cctx

    def _loss_terms(self, logits, kind, slot, losses, accumulate, want_grad):
        dy = None
        y = logits
        if want_grad:
            dy = Act.empty(y.n, y.t, y.h, y.w, y.c, y.dtype)
        ops.loss_fwd_bwd_act(y, kind, losses[slot:slot + 1], accumulate, dy, 1.0)
        return dy

    # ------------------------------------------------------------------ input prefetch
    def prefetch(self, xc_host, xg_host):
        """Start the host->device copy of a FUTURE batch (pinned host tensors) on a copy stream, so that it travels over
        PCIe while the current iteration computes; `train_step` called later with the same two host tensors picks the
        staged copy up instead of issuing a blocking H2D at its start.  The reference moves each batch with
        `.to(device, non_blocking=True)` right before use (trainer.py:293-297), which serialises copy and compute.
        Two staging slots are used alternately; a slot is not overwritten before the step that consumed it has read it."""
        if xc_host.is_cuda:
            return
        self._prepare()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = [None, None]
            self._stage_i = 0
        i = self._stage_i
        self._stage_i ^= 1
        slot = self._stage[i]
        if (slot is None or slot["xc"].shape != xc_host.shape or slot["xg"].shape != xg_host.shape
                or slot["xc"].dtype != xc_host.dtype or slot["xg"].dtype != xg_host.dtype):
            slot = self._stage[i] = {"xc": torch.empty(xc_host.shape, dtype=xc_host.dtype, device=self.device),
                                     "xg": torch.empty(xg_host.shape, dtype=xg_host.dtype, device=self.device),
                                     "ready": torch.cuda.Event(), "consumed": None}
        if slot["consumed"] is not None:
            self._copy_stream.wait_event(slot["consumed"])
        with torch.cuda.stream(self._copy_stream):
            slot["xc"].copy_(xc_host, non_blocking=True)
            slot["xg"].copy_(xg_host, non_blocking=True)
            slot["ready"].record(self._copy_stream)
        slot["key"] = (weakref.ref(xc_host), weakref.ref(xg_host))      # matched by object identity, not by address

    def _take_prefetched(self, xc_real, xg_real):
        """(device xc, device xg, slot) if this host batch was staged by prefetch(), else None"""
        if xc_real.is_cuda or self._copy_stream is None:
            return None
        for slot in self._stage:
            key = slot.get("key") if slot is not None else None
            if key is not None and key[0]() is xc_real and key[1]() is xg_real:
                slot["key"] = None
                torch.cuda.current_stream().wait_event(slot["ready"])
                return slot
        return None

    def train_step(self, xc_real, xg_real, t_rand=None):
        """One iteration of trainer.py:279-363.  xc_real (B,3,T,64,64), xg_real (B,C,T,64,64) on the device.
        Returns a device tensor [loss_idis, loss_vdis, loss_gdis, loss_gen].

        With the device RNG the whole iteration (~500 kernel launches, every one of them static in shape and
        address) is captured once per update pattern into a CUDA graph and replayed; the frame index t_rand and
        the Adam step counters live in device memory so nothing has to be re-recorded."""
        self._prepare()
        self._check_bound()
        if t_rand is None:
            t_rand = np.random.randint(self.models["ggen"].video_length)                        # trainer.py:279
        if self.use_cuda_graph and engine.rng().mode in ("device", "staged"):
            return self._graph_step(xc_real, xg_real, int(t_rand))
        return self._eager_step(xc_real, xg_real, t_rand)

    def _eager_step(self, xc_real, xg_real, t_rand):
        if not xc_real.is_cuda:      # host (pinned) batch: H2D copy as at trainer.py:293-297 (or the copy staged by prefetch())
            staged = self._take_prefetched(xc_real, xg_real)
            if staged is not None:
                xc_real, xg_real = staged["xc"].clone(), staged["xg"].clone()
                staged["consumed"] = torch.cuda.Event()
                staged["consumed"].record(torch.cuda.current_stream())
            else:
                xc_real = xc_real.to(self.device, non_blocking=True)
                xg_real = xg_real.to(self.device, non_blocking=True)
        # packed-weight cache owned by this trainer (keys are ids of its parameters); it only lives WITHIN one iteration
        # (D: real + fake passes, G: forward + backward) - weights may change between iterations without the cache
        # knowing (graph replays, load_state_dict, a hook that edits parameters), so every iteration starts empty
        self._wcache.clear()
        engine.WCACHE = self._wcache
        engine.PACK_GROUP = self._pack_group  # a miss re-packs every logged weight of the same network in one launch
        engine.PACK_LOG = self._pack_log
        if self._zero_pool is not None:
            self._zero_pool.begin_step()
        ops.ZERO_POOL = self._zero_pool       # zero-padded scratch buffers reused across iterations (ops.ZeroPool)
        ops.STRICT_TC = self._strict_tc
        try:
            return self._train_step(xc_real, xg_real, t_rand)
        finally:
            engine.WCACHE = None
            engine.PACK_GROUP = None
            engine.PACK_LOG = None
            ops.ZERO_POOL = None
            ops.STRICT_TC = False

    GRAPH_WARMUP = 2   # eager iterations per update pattern before capture (lazy module loads, smem opt-in, NCCL setup)

    def _graph_step(self, xc_real, xg_real, t_rand):
        cfg = self.configs
        ggen, cgen = self.models["ggen"], self.models["cgen"]
        key = (self.iteration % cfg["num_gen_update"] == 0, self.iteration % cfg["num_dis_update"] == 0,
               ggen.training, cgen.training, tuple(xc_real.shape), tuple(xg_real.shape), xc_real.dtype, xg_real.dtype)
        slot = self._graphs.setdefault(key, [0, None, None, None])
        if slot[1] is None and slot[0] < self.GRAPH_WARMUP:
            slot[0] += 1
            return self._eager_step(xc_real, xg_real, t_rand)
        if slot[3] is None:   # static inputs of this graph: the real batch and the frame index
            slot[3] = (torch.empty(xc_real.shape, dtype=xc_real.dtype, device=self.device),
                       torch.empty(xg_real.shape, dtype=xg_real.dtype, device=self.device),
                       torch.zeros(1, dtype=torch.int32, device=self.device))
        gxc, gxg, t_dev = slot[3]
        staged = self._take_prefetched(xc_real, xg_real)
        if staged is not None:            # device-to-device from the staging slot (the PCIe copy already happened)
            gxc.copy_(staged["xc"], non_blocking=True)
            gxg.copy_(staged["xg"], non_blocking=True)
            staged["consumed"] = torch.cuda.Event()
            staged["consumed"].record(torch.cuda.current_stream())
        else:
            gxc.copy_(xc_real, non_blocking=True)
            gxg.copy_(xg_real, non_blocking=True)
        t_dev.fill_(t_rand)     # the value travels as a kernel argument, ordered on the stream (no reused pinned scalar)
        if slot[1] is None:
            # Capture.  The generators' train/eval flags change inside the step (trainer.py:338-339); restore them so the
            # recorded pattern matches `key`, then let the first replay perform the iteration.
            modes = (ggen.training, cgen.training)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.lib().dcv_launch_count()
            with torch.cuda.graph(graph):
                losses = self._eager_step(gxc, gxg, t_dev)
            slot.append(_lib.lib().dcv_launch_count() - l0)      # kernels of this library recorded in the graph
            ggen.train(modes[0])
            cgen.train(modes[1])
            self._wcache.clear()                 # packed weights created during capture live in the graph's pool
            slot[1], slot[2] = graph, losses
        slot[1].replay()
        self.replayed_launches += slot[4]
        # The replay ran the Adam kernels on the device: every packed bf16 weight the HOST-side cache still holds (left
        # by an eager warm-up iteration of another graph key) is stale now and must not be reused by a later eager step.
        self._wcache.clear()
        ggen.train()                             # host-side flags the replay cannot set (trainer.py:338-339)
        cgen.train()
        return slot[2].clone()

    def _train_step(self, xc_real, xg_real, t_rand):
        cfg = self.configs
        B = cfg["batchsize"]
        ggen, cgen = self.models["ggen"], self.models["cgen"]
        T = ggen.video_length
        losses = torch.zeros(4, dtype=torch.float32, device=self.device)
        slot = {"idis": 0, "vdis": 1, "gdis": 2}
        L = self.loss

        # ---------------- discriminator phase (trainer.py:285-333)
        for n in self._dnames:
            self.models[n].train()
        upd_d = self.iteration % cfg["num_gen_update"] == 0                                     # sic, trainer.py:318
        xc_r, xg_r = self._to_clip(xc_real), self._to_clip(xg_real, self.models["ggen"].channel)
        def spec(kind, slot_of, acc_of, want_grad, names):
            return {n: {"kind": kind, "out": losses[slot_of(n):slot_of(n) + 1], "accumulate": acc_of(n), "want_grad": want_grad} for n in names}
        real = self._dis_forward(xg_r, xc_r, t_rand, upd_d, spec(L.KIND_REAL, slot.get, lambda n: False, upd_d, self._dnames))
        xg_f, xc_f, _, _ = self._generate(B, ggen.training, cgen.training, save=False)
        fake = self._dis_forward(xg_f.reshape_nt(B, T), xc_f.reshape_nt(B, T), t_rand, upd_d,
                                 spec(L.KIND_FAKE, slot.get, lambda n: True, upd_d, self._dnames))
        grads = {}
        for n in self._dnames:     # heads that could not fuse their loss term (fp32 / CUDA-core path): the separate loss kernel
            dr = real[n][2]["dlogits"] if real[n][2]["loss_fused"] else self._loss_terms(real[n][0], L.KIND_REAL, slot[n], losses, False, upd_d)
            df = fake[n][2]["dlogits"] if fake[n][2]["loss_fused"] else self._loss_terms(fake[n][0], L.KIND_FAKE, slot[n], losses, True, upd_d)
            grads[n] = (dr, df)
        if upd_d:
            for n in self._dnames:
                sink = engine.GradSink(self._flat[n].grad_of)
                self._plans[n].backward(real[n][1], grads[n][0], sink, need_dx=False)
                self._plans[n].backward(fake[n][1], grads[n][1], sink, need_dx=False)
                self._reduce_launch(n)                  # this network's bucket travels while the next one's backward runs
        del real, fake, grads

        # ---------------- generator phase (trainer.py:338-368)
        ggen.train()
        cgen.train()
        upd_g = self.iteration % cfg["num_dis_update"] == 0                                     # sic, trainer.py:355
        # The generators' forward pass does not read the discriminators: it runs BEFORE the discriminators' Adam steps so
        # that their gradient buckets cross NVLink underneath it (data-parallel runs); the values are those of the
        # reference's order (opt_*dis.step() consume no random numbers and do not touch the generators).
        xg_f, xc_f, gctx, cctx = self._generate(B, True, True, save=upd_g)
        if upd_d:
            self._reduce_wait("d")
            for n in self._dnames:
                self._flat[n].adam_step(1.0 / self.world)
        xg_clip, xc_clip = xg_f.reshape_nt(B, T), xc_f.reshape_nt(B, T)
        gen_terms = ["idis", "vdis"] + (["gdis"] if (self.use_gdis and L.gen_uses_gdis) else [])
        fake = self._dis_forward(xg_clip, xc_clip, t_rand, upd_g, spec(L.KIND_GEN, lambda n: 3, lambda n: n != "idis", upd_g, gen_terms))
        dls = {}
        for i, n in enumerate(gen_terms):
            dls[n] = fake[n][2]["dlogits"] if fake[n][2]["loss_fused"] else self._loss_terms(fake[n][0], L.KIND_GEN, 3, losses, i > 0, upd_g)
        if upd_g:
            none = engine.GradSink()
            dxg = Act.empty(B, T, 64, 64, xg_clip.c, self.dtype)
            dxc = Act.empty(B, T, 64, 64, 3, self.dtype)
            g_v, c_v = self._plans["vdis"].backward(fake["vdis"][1], dls["vdis"], none, need_dx=True, need_dw=False)
            ops.copy_cl(g_v, dxg)
            ops.copy_cl(c_v, dxc)
            g_i, c_i = self._plans["idis"].backward(fake["idis"][1], dls["idis"], none, need_dx=True, need_dw=False)
            ops.frame_scatter(dxg, t_rand, g_i, True)
            ops.frame_scatter(dxc, t_rand, c_i, True)
            if "gdis" in dls:
                g_g, _ = self._plans["gdis"].backward(fake["gdis"][1], dls["gdis"], none, need_dx=True, need_dw=False)
                ops.axpy(g_g, dxg, True)
            sink_c = engine.GradSink(self._flat["cgen"].grad_of)
            dxg_c = self._plans["cgen"].backward(cctx, dxc.reshape_nt(B * T, 1), sink_c, need_dx=True)
            self._reduce_launch("cgen")                 # 10.6 M floats, reduced under ggen's backward
            if dxg_c is not None:
                ops.axpy(dxg_c.reshape_nt(B, T), dxg, True)
            sink_g = engine.GradSink(self._flat["ggen"].grad_of)
            self._plans["ggen"].backward(gctx, dxg.reshape_nt(B * T, 1), sink_g)
            self._reduce_launch("ggen")
            self._reduce_wait("g")
            self._flat["ggen"].adam_step(1.0 / self.world)
            self._flat["cgen"].adam_step(1.0 / self.world)
            self._flat["ggen"].adam_step(1.0 / self.world)                                      # trainer.py:357-359
        return losses

    # ------------------------------------------------------------------ logging of device-side losses
    def sync_optimizer_state(self):
        if self._flat:
            for f in self._flat.values():
                f.sync_optimizer_state()

    def _flush_losses(self):
        if not self._pending:
            return
        vals = torch.stack([l for _, _, l in self._pending]).cpu().tolist()
        for (it, ep, _), v in zip(self._pending, vals):
            self.logger.update("iteration", it)
            self.logger.update("epoch", ep)
            self.logger.update("loss_idis", v[0])
            self.logger.update("loss_vdis", v[1])
            if self.use_gdis:
                self.logger.update("loss_gdis", v[2])
            self.logger.update("loss_gen", v[3])
        self._pending = []

    def train(self):
        """Start training (trainer.py:226-392)."""
        self._prepare()
        ggen, cgen = self.models["ggen"], self.models["cgen"]
        for name in ("loss_gen", "loss_idis", "loss_vdis", "loss_gdis"):
            self.logger.define(name, MetricType.Loss)
        for m in self.configs["evaluation"]["metrics"]:
            self.logger.define(m, MetricType.Float)
        self.log_hparams()
        self.log_samples(ggen, cgen, 0)
        self.evaluate(ggen, cgen)
        if hasattr(self.logger, "print_header"):
            self.logger.print_header()
        cfg = self.configs
        for _ in range(cfg["n_epochs"]):
            self.epoch += 1
            it = iter(self.dataloader)
            batch = next(it, None)
            while batch is not None:
                self.iteration += 1
                xc_real, xg_real = batch["color"], batch[self.geometric_info]
                losses = self.train_step(xc_real, xg_real)
                batch = next(it, None)
                if batch is not None and not batch["color"].is_cuda and batch["color"].is_pinned():
                    self.prefetch(batch["color"], batch[self.geometric_info])   # next batch crosses PCIe under this iteration
                self._pending.append((self.iteration, self.epoch, losses))
                if self.iteration % cfg["snapshot_interval"] == 0:
                    self.save_params()
                if self.iteration % cfg["log_samples_interval"] == 0:
                    self.log_samples(ggen, cgen, self.iteration)
                if self.iteration % cfg["evaluation_interval"] == 0:
                    self._flush_losses()
                    self.evaluate(ggen, cgen)
                if self.iteration % cfg["log_interval"] == 0:
                    self._flush_losses()
                    self.logger.log()
                    self.logger.clear()
        self._flush_losses()
        self.sync_optimizer_state()
        self.save_params()
        self.log_samples(ggen, cgen, self.iteration)
