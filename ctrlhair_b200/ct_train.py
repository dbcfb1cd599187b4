"""Colour/texture training step (config 045) on the B200, behind the reference Solver's surface.

Mirrors color_texture_branch/solver.py `Solver` (forward / forward_d / forward_g, :85-245), the Adam optimizers of
:52-55, `train()` of my_torchlib/train_utils.py:54-89 and the loop body of color_texture_branch/train.py:115-148:

    solver = SolverB200(cfg, device, local_rank)          # cfg: dict/Namespace with the lambda_* / lr / beta fields
    solver.load_state_dicts(G_sd, D_sd, rgb_predictor_sd, curliness_predictor_sd)
    for i in range(2):
        data = ...                                         # the dict train.py:118-126 builds
        loss_dict = {}
        solver.forward(data)
        if i == 0: solver.forward_d(loss_dict); train(cfg, loss_dict, optimizers=[solver.D_optimizer])
        else:      solver.forward_g(loss_dict); train(cfg, loss_dict, optimizers=[solver.G_optimizer])

`forward` draws the reference's random numbers from the same generators in the same order (three in-place
`random.shuffle`s + one `random.random()`; `torch.rand(B, 1)` for alpha_gp in forward_d), so a run seeded like the
reference takes the same permutations.  forward_d / forward_g launch forward + losses + backward of the sub-step as
ONE CUDA graph; `optimizer.step()` all-reduces the flat gradient buffer (world size > 1) and runs the Adam kernel.
There is no CPU path: without the CUDA library or a CC 10.x device construction raises.
"""
import ctypes as C
import random

import torch

from . import _lib, parallel
from ._lib import CTT_D, CTT_FROZEN, CTT_G, CTT_LOSS_NAMES, CtTrainBatch, CtTrainConfig

BN_EPS = 1e-5
DEFAULTS = {  # config 045 after LossUpdater.update(0): color_texture_branch/config.py:16-39 + defaults :52-96
    "lambda_adv": 1.0, "lambda_gp": 10.0, "lambda_info": 1.0, "lambda_info_curliness": 1.0, "lambda_rec": 1000.0,
    "lambda_rgb": 0.01, "lambda_pca_std": 0.01, "lambda_moment_1": 0.01, "lambda_moment_2": 0.01,
    "lambda_cls_curliness": 0.1, "lambda_orthogonal": 0.1, "lr_d": 2e-4, "lr_g": 2e-4, "beta1": 0.5, "beta2": 0.999,
    "gan_input_from_encoder_prob": 0.3,
}
D_LOSSES = ("lambda_adv", "lambda_gp", "lambda_info", "lambda_rec", "lambda_moment_1", "lambda_moment_2",
            "lambda_info_curliness")
G_LOSSES = ("lambda_adv", "lambda_info", "lambda_rec", "lambda_rgb", "lambda_pca_std", "lambda_info_curliness",
            "lambda_cls_curliness", "lambda_orthogonal")


def _cfg_get(cfg, key):
    if cfg is None:
        return DEFAULTS[key]
    v = cfg.get(key, None) if isinstance(cfg, dict) else getattr(cfg, key, None)
    if isinstance(v, dict):   # step schedule {start_step: weight}: the LossUpdater resolves it; take step 0
        v = v[min(v)]
    return DEFAULTS[key] if v is None or v == {} else float(v)


def fold_predictor(sd, n_hidden=3):
    """Eval-mode BatchNorm1d folded into the preceding Linear (LinearBlock, my_torchlib/module.py:56-64)."""
    out = {}
    for i in range(n_hidden + 1):
        w, b = sd["net.%d.fc.weight" % i].double(), sd["net.%d.fc.bias" % i].double()
        if i < n_hidden:
            s = sd["net.%d.norm.weight" % i].double() / torch.sqrt(sd["net.%d.norm.running_var" % i].double() + BN_EPS)
            t = sd["net.%d.norm.bias" % i].double() - sd["net.%d.norm.running_mean" % i].double() * s
            w, b = w * s[:, None], b * s + t
        out["net.%d.fc.weight" % i], out["net.%d.fc.bias" % i] = w.float(), b.float()
    return out


class _FusedAdam:
    """Stands where torch.optim.Adam stands in train(): zero_grad() is part of the sub-step graph, step() = gradient
    all-reduce (mean over ranks, as DDP) + the Adam kernel."""

    def __init__(self, solver, which):
        self.solver, self.which = solver, which

    def zero_grad(self):
        pass

    def step(self):
        s = self.solver
        caller = torch.cuda.current_stream(s.device)
        s.stream.wait_stream(caller)  # anything the caller queued (e.g. a state_dict load) is ordered before the step
        with torch.cuda.device(s.device), torch.cuda.stream(s.stream):
            if s.world > 1:
                parallel.allreduce_mean_(s.region(1, self.which))
            _lib.check(s.lib.chb_cttrain_adam(s.handle, self.which, C.c_void_p(s.stream.cuda_stream)))
        caller.wait_stream(s.stream)  # ... and the updated parameters are visible to what the caller does next


def train(cfg, loss_dict, optimizers, step=0, writer=None, flag="", retain_graph=False, write_log=False):
    """my_torchlib/train_utils.py:54-89 on the fused path: the weighted sum and the backward already happened inside
    forward_d / forward_g; what is left is the optimizer step.  No host sync (the reference's per-term NaN checks
    cost >= 10 syncs per step; `loss_dict['total']` can be checked by the caller when it wants to)."""
    if len(loss_dict) == 0:
        return
    for o in optimizers:
        o.zero_grad()
        o.step()


class SolverB200:
    def __init__(self, cfg=None, device="cuda", local_rank=-1, training=True, batch_size=None, use_graph="graph"):
        """use_graph: how a sub-step reaches the GPU — "graph" (default; True means the same): an explicit CUDA graph, one
        node per operation with an edge for every real data dependency (independent GEMMs run side by side: 0.58 ms per
        train.py iteration at B = 32); "persistent": ONE cooperative kernel walking the operation list with grid barriers
        between dependent operations (1.1 ms: a grid barrier plus a serialised tile per phase costs more than a graph
        edge); False / 0: plain launches (1 linear chain).  All three run the same arithmetic."""
        self.lib = _lib.load()
        self.device = torch.device(device if local_rank < 0 else "cuda:%d" % local_rank)
        if self.device.type != "cuda":
            raise _lib.ChbError("SolverB200 needs a CUDA device (there is no CPU path)")
        if batch_size is None:
            batch_size = int(cfg["batch_size"] if isinstance(cfg, dict) else cfg.batch_size)
        self.cfg, self.B = cfg, batch_size
        self.enc_prob = _cfg_get(cfg, "gan_input_from_encoder_prob")
        c = CtTrainConfig()
        c.batch = batch_size
        for k in DEFAULTS:
            if k.startswith("lambda_"):
                setattr(c, k, _cfg_get(cfg, k))
        # the reference wires lr_d to G's optimizer and lr_g to D's (solver.py:52-55); both are 2e-4 in every shipped
        # config.  The fused Adam kernel takes ONE learning rate: differing values are rejected instead of ignored.
        lr_d, lr_g = _cfg_get(cfg, "lr_d"), _cfg_get(cfg, "lr_g")
        if lr_d != lr_g:
            raise _lib.ChbError("SolverB200: lr_d (%g) != lr_g (%g) is not supported by the fused Adam step" % (lr_d, lr_g))
        # lambda_rec_img (045: 0 -> 1000 at step 600 000) needs no guard: forward_rec_img runs under no_grad
        # (pix2pix_model.py:60), so it changes the logged loss only, never a gradient.  lambda_adv_noise would add a
        # third network (Model_D_noise) that this fused step does not hold: reject it.
        v = None if cfg is None else (cfg.get("lambda_adv_noise", None) if isinstance(cfg, dict)
                                      else getattr(cfg, "lambda_adv_noise", None))
        vals = list(v.values()) if isinstance(v, dict) else [v]
        if any(x not in (None, {}, 0, 0.0) for x in vals):
            raise _lib.ChbError("SolverB200: lambda_adv_noise != 0 (noise discriminator) is not implemented")
        c.lr, c.beta1, c.beta2, c.eps = lr_d, _cfg_get(cfg, "beta1"), _cfg_get(cfg, "beta2"), 1e-8
        c.use_graph = {"persistent": 2, "graph": 1, True: 1, False: 0, None: 0}.get(use_graph, use_graph)
        if c.use_graph not in (0, 1, 2):
            raise _lib.ChbError("use_graph must be 'persistent', 'graph', True/False or 0/1/2")
        self.mode = c.use_graph
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_cttrain_create(C.byref(c), C.byref(self.handle)))
            self.state = torch.zeros(self.lib.chb_cttrain_state_floats(self.handle), dtype=torch.float32,
                                     device=self.device)
            self.workspace = torch.zeros(self.lib.chb_cttrain_workspace_bytes(self.handle), dtype=torch.uint8,
                                         device=self.device)
            _lib.check(self.lib.chb_cttrain_bind(self.handle, C.c_void_p(self.state.data_ptr()),
                                                 C.c_void_p(self.workspace.data_ptr())))
            self.stream = torch.cuda.Stream(device=self.device)
        self.table = {}
        name = C.create_string_buffer(128)
        off, num, grp = C.c_int64(), C.c_int64(), C.c_int()
        for i in range(self.lib.chb_cttrain_num_tensors(self.handle)):
            _lib.check(self.lib.chb_cttrain_tensor_info(self.handle, i, name, 128, C.byref(off), C.byref(num),
                                                        C.byref(grp)))
            self.table[name.value.decode()] = (off.value, num.value, grp.value)
        self.losses = torch.zeros(len(CTT_LOSS_NAMES), dtype=torch.float32, device=self.device)
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.D_optimizer, self.G_optimizer = _FusedAdam(self, CTT_D), _FusedAdam(self, CTT_G)
        self._data = None
        self._keep = None

    def __del__(self):
        # (at interpreter shutdown torch.nn may already be torn down: bypass nn.Module.__setattr__, never raise)
        h = self.__dict__.get("handle")
        if h:
            self.__dict__["handle"] = None
            try:
                self.lib.chb_cttrain_destroy(h)
            except Exception:
                pass

    # ---- state ---------------------------------------------------------------------------------------------
    def region(self, region, group):
        off, num = C.c_int64(), C.c_int64()
        _lib.check(self.lib.chb_cttrain_region(self.handle, region, group, C.byref(off), C.byref(num)))
        return self.state[off.value:off.value + num.value]

    def _net_view(self, prefix, region):
        group = {"D.": CTT_D, "G.": CTT_G}.get(prefix, CTT_FROZEN)
        base = self.region(region, group)
        return {k[len(prefix):]: (base, o, n) for k, (o, n, g) in self.table.items() if k.startswith(prefix)}

    def _load(self, prefix, sd, strict=True):
        view = self._net_view(prefix, 0)
        sd = {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}
        if strict and set(sd) != set(view):
            raise KeyError("state_dict keys do not match %s: missing %s, unexpected %s" %
                           (prefix, sorted(set(view) - set(sd)), sorted(set(sd) - set(view))))
        for k, (base, o, n) in view.items():
            t = sd[k].detach().to(torch.float32).reshape(-1)
            if t.numel() != n:
                raise ValueError("%s%s has %d elements, expected %d" % (prefix, k, t.numel(), n))
            base[o:o + n].copy_(t)

    def load_state_dicts(self, gen_sd, dis_sd, rgb_predictor_sd, curliness_predictor_sd):
        """ckpt['Model_G'], ckpt['Model_D'] (strict keys, hair_editor.py:70-71) and the two frozen Predictor dicts."""
        self._load("G.", gen_sd)
        self._load("D.", dis_sd)
        self._load("P.", fold_predictor(rgb_predictor_sd))
        self._load("C.", fold_predictor(curliness_predictor_sd))
        torch.cuda.synchronize(self.device)

    def _export(self, prefix, region, ref_sd=None):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        out = {}
        for k, (base, o, n) in self._net_view(prefix, region).items():
            t = base[o:o + n].clone()
            out[k] = t.reshape(ref_sd[k].shape) if ref_sd is not None else t
        return out

    def gen_state_dict(self, like=None):
        return self._export("G.", 0, like)

    def dis_state_dict(self, like=None):
        return self._export("D.", 0, like)

    def gen_grads(self, like=None):
        return self._export("G.", 1, like)

    def dis_grads(self, like=None):
        return self._export("D.", 1, like)

    # ---- the reference surface -----------------------------------------------------------------------------
    def draw_randomness(self):
        """solver.py:98-111 (python `random`): three successive in-place shuffles of one index list, then the coin."""
        lst = list(range(self.B))
        random.shuffle(lst)
        p1 = list(lst)
        random.shuffle(lst)
        p2 = list(lst)
        random.shuffle(lst)
        p3 = list(lst)
        use_enc = bool(self.enc_prob) and random.random() < self.enc_prob
        return {"p1": p1, "p2": p2, "p3": p3, "use_enc": use_enc}

    def forward(self, data, randomness=None):
        for k in ("code", "rgb_mean", "pca_std", "noise", "noise_curliness", "curliness_label"):
            if data[k].shape[0] != self.B:
                raise _lib.ChbError("data['%s'] has %d rows, solver was built for batch %d" % (k, data[k].shape[0], self.B))
        self._data = data
        self._rnd = randomness if randomness is not None else self.draw_randomness()

    def _run(self, which, alpha=None):
        if self._data is None:
            raise _lib.ChbError("call forward(data) first")
        d, r = self._data, self._rnd
        dev = self.device
        f32 = lambda t: t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()  # noqa: E731
        i32 = lambda p: torch.as_tensor(p, dtype=torch.int32).to(dev, non_blocking=True)  # noqa: E731
        caller = torch.cuda.current_stream(dev)
        # the step runs on a private stream (graph replay): order it after the caller's stream, on which device-resident
        # batch tensors may still be being produced, and make the caller's stream wait for the losses afterwards so that
        # a reference-style `loss_dict[k].item()` / NaN check never reads an unwritten buffer
        self.stream.wait_stream(caller)
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            keep = {k: f32(d[k]) for k in ("code", "rgb_mean", "pca_std", "noise", "noise_curliness", "curliness_label")}
            keep["p1"], keep["p2"], keep["p3"] = i32(r["p1"]), i32(r["p2"]), i32(r["p3"])
            b = CtTrainBatch()
            for k in ("code", "rgb_mean", "pca_std", "noise", "noise_curliness", "curliness_label"):
                setattr(b, k, keep[k].data_ptr())
            b.perm_rgb, b.perm_curliness, b.perm_noise = keep["p1"].data_ptr(), keep["p2"].data_ptr(), keep["p3"].data_ptr()
            if alpha is not None:
                keep["alpha"] = f32(alpha)
                b.alpha_gp = keep["alpha"].data_ptr()
            b.noise_from_encoder = 1 if r["use_enc"] else 0
            losses = torch.empty(len(CTT_LOSS_NAMES), dtype=torch.float32, device=dev)
            _lib.check(self.lib.chb_cttrain_step(self.handle, which, C.byref(b), C.c_void_p(losses.data_ptr()),
                                                 C.c_void_p(self.stream.cuda_stream)))
            for t in keep.values():
                t.record_stream(self.stream)
            for k in ("code", "rgb_mean", "pca_std", "noise", "noise_curliness", "curliness_label"):
                if isinstance(d[k], torch.Tensor) and d[k].is_cuda:
                    d[k].record_stream(self.stream)  # inputs already on the device pass through f32() as views
        caller.wait_stream(self.stream)
        losses.record_stream(caller)
        self._keep = keep
        self.losses = losses
        return losses

    def forward_d(self, loss_dict, alpha_gp=None):
        """solver.py:218-245; alpha_gp defaults to torch.rand(B, 1) on the CPU generator like solver.py:199."""
        if alpha_gp is None:
            alpha_gp = torch.rand(self.B, 1)
        losses = self._run(CTT_D, alpha_gp)
        for k in D_LOSSES + ("total",):
            loss_dict[k] = losses[CTT_LOSS_NAMES.index(k)]

    def forward_g(self, loss_dict):
        """solver.py:119-166."""
        losses = self._run(CTT_G)
        for k in G_LOSSES + ("total",):
            loss_dict[k] = losses[CTT_LOSS_NAMES.index(k)]

    def synchronize(self):
        self.stream.synchronize()

    def launches(self, which):
        return self.lib.chb_cttrain_launches(self.handle, which)

    def schedule(self, which):
        """After the first use of a sub-step: (operations, dependency edges) of its graph, or (operations, grid barriers)
        of its persistent kernel."""
        n, b = C.c_int(), C.c_int()
        _lib.check(self.lib.chb_cttrain_schedule(self.handle, which, C.byref(n), C.byref(b)))
        return n.value, b.value
