"""CPU oracle for the DCVGAN training step  --  TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A functional (no nn.Module) restatement, in plain PyTorch fp32 on the CPU, of the arithmetic the
reference performs on its hot path.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import this file; dcvgan_b200/ never does.

Parity status: PINNED.  tests/test_oracle.py checks every function below against golden vectors
produced by the unmodified reference modules and the unmodified reference Trainer.train() loop
(oracle/make_golden.py, run in the build container where /root/reference is mounted; fixtures in
tests/golden/).

Third-party arithmetic: all of it lives in PyTorch (the reference pins torch==1.2.0,
requirements.txt:15; this image has 2.11).  conv / conv_transpose / batch_norm / dropout2d / GRU
gate equations / BCE-with-logits / Adam are used through torch.nn.functional / torch.optim exactly
where the reference uses the corresponding nn.Module, so the semantics are the library's own.

Parameter containers are dicts keyed by the reference's state_dict keys.
Every function cites the reference lines it follows (paths relative to the reference's src/).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS, BN_MOM = 1e-5, 0.1


# ------------------------------------------------------------------------------------------ config
def normalise_config(cfg):
    """Bring the legacy YAML schemas to the one train.py reads at HEAD.

    gen -> ggen/cgen, geometric_info str -> {name, channel}, default loss / update ratios /
    evaluation block; a missing gdis block means "gradient discriminator disabled".
    """
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
    if "gen" in cfg:
        g = cfg.pop("gen")
        opt = g.get("optimizer", {"lr": 0.0002, "decay": 0.00001})
        cfg.setdefault("ggen", {"dim_z_content": g["dim_z_content"], "dim_z_motion": g["dim_z_motion"],
                                "ngf": g["ngf"], "optimizer": dict(opt)})
        cfg.setdefault("cgen", {"dim_z_color": g["dim_z_color"], "ngf": g["ngf"], "optimizer": dict(opt)})
    gi = cfg.get("geometric_info", "depth")
    if not isinstance(gi, dict):
        name = gi
        cfg["geometric_info"] = {"name": name, "channel": {"depth": 1, "optical-flow": 2, "segmentation": 25}[name]}
    cfg.setdefault("loss", "adversarial-loss")
    cfg.setdefault("num_gen_update", 1)
    cfg.setdefault("num_dis_update", 1)
    cfg.setdefault("evaluation", {"batchsize": 50, "num_samples": 0, "metrics": []})
    if "gdis" not in cfg:
        cfg["gdis"] = {"use_noise": False, "noise_sigma": 0.2, "ndf": 32,
                       "optimizer": {"lr": 0.0002, "decay": 0.00001}, "enabled": False}
    else:
        cfg["gdis"].setdefault("enabled", True)
    return cfg


# ------------------------------------------------------------------------------------------ init
def _conv_default(shape, gen):
    """torch's default Conv/ConvTranspose init: kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(fan_in), +)"""
    fan_in = shape[1] * int(np.prod(shape[2:]))
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def _normal(shape, mean, std, gen):
    return torch.randn(shape, generator=gen) * std + mean


def _bn(prefix, c, P, gen, two_d):
    # util.py:186-195: BatchNorm2d weight N(1, 0.02), bias 0; BatchNorm3d keeps torch defaults (1, 0)
    P[prefix + ".weight"] = _normal((c,), 1.0, 0.02, gen) if two_d else torch.ones(c)
    P[prefix + ".bias"] = torch.zeros(c)
    P[prefix + ".running_mean"] = torch.zeros(c)
    P[prefix + ".running_var"] = torch.ones(c)
    P[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def init_ggen(cfg, gen):
    """generator.py:58-80 layer shapes; util.py:186-195 init (N(0,0.02) for 2-D convs)."""
    g, C, ngf = cfg["ggen"], cfg["geometric_info"]["channel"], cfg["ggen"]["ngf"]
    dm, dz = g["dim_z_motion"], g["dim_z_motion"] + g["dim_z_content"]
    P = {}
    k = 1.0 / math.sqrt(dm)
    for name, shape in (("weight_ih", (3 * dm, dm)), ("weight_hh", (3 * dm, dm)), ("bias_ih", (3 * dm,)), ("bias_hh", (3 * dm,))):
        P["recurrent." + name] = (torch.rand(shape, generator=gen) * 2 - 1) * k
    chans = [dz, ngf * 8, ngf * 4, ngf * 2, ngf, C]
    for i in range(5):
        P[f"main.{3 * i}.weight"] = _normal((chans[i], chans[i + 1], 4, 4), 0.0, 0.02, gen)
        if i < 4:
            _bn(f"main.{3 * i + 1}", chans[i + 1], P, gen, True)
    return P


def init_cgen(cfg, gen):
    """generator.py:171-176,203-207,238-248,272-277,323-344."""
    C, ngf, dz = cfg["geometric_info"]["channel"], cfg["cgen"]["ngf"], cfg["cgen"]["dim_z_color"]
    P = {"inconv.main.0.weight": _normal((ngf, C, 3, 3), 0.0, 0.02, gen)}
    down = [(ngf, ngf), (ngf, 2 * ngf), (2 * ngf, 4 * ngf), (4 * ngf, 4 * ngf), (4 * ngf, 4 * ngf), (4 * ngf, 4 * ngf)]
    for i, (ci, co) in enumerate(down):
        P[f"down_blocks.{i}.main.0.weight"] = _normal((co, ci, 4, 4), 0.0, 0.02, gen)
        _bn(f"down_blocks.{i}.main.1", co, P, gen, True)
    up = [(4 * ngf + dz, 4 * ngf), (8 * ngf, 4 * ngf), (8 * ngf, 4 * ngf), (8 * ngf, 2 * ngf), (4 * ngf, ngf), (2 * ngf, ngf)]
    for i, (ci, co) in enumerate(up):
        P[f"up_blocks.{i}.main.0.weight"] = _normal((ci, co, 4, 4), 0.0, 0.02, gen)
        _bn(f"up_blocks.{i}.main.1", co, P, gen, True)
    P["outconv.main.0.weight"] = _normal((2 * ngf, 3, 3, 3), 0.0, 0.02, gen)
    return P


def init_dis(kind, cfg, gen):
    """discriminator.py:79-102 (idis), :180-207 (vdis), :285-306 (gdis).  Conv3d / BatchNorm3d are not in
    util.init_weights' type list, so they keep torch's defaults."""
    C, ndf = cfg["geometric_info"]["channel"], cfg[kind]["ndf"]
    P = {}
    if kind == "idis":
        P["conv_g.1.weight"] = _normal((ndf // 2, C, 4, 4), 0.0, 0.02, gen)
        P["conv_c.1.weight"] = _normal((ndf // 2, 3, 4, 4), 0.0, 0.02, gen)
        P["main.1.weight"] = _normal((ndf * 2, ndf, 4, 4), 0.0, 0.02, gen)
        _bn("main.2", ndf * 2, P, gen, True)
        P["main.5.weight"] = _normal((ndf * 4, ndf * 2, 4, 4), 0.0, 0.02, gen)
        _bn("main.6", ndf * 4, P, gen, True)
        P["main.9.weight"] = _normal((1, ndf * 4, 4, 4), 0.0, 0.02, gen)
    elif kind == "vdis":
        P["conv_g.0.weight"] = _conv_default((ndf // 2, C, 4, 4, 4), gen)
        P["conv_c.0.weight"] = _conv_default((ndf // 2, 3, 4, 4, 4), gen)
        P["main.1.weight"] = _conv_default((ndf * 2, ndf, 4, 4, 4), gen)
        _bn("main.2", ndf * 2, P, gen, False)
        P["main.5.weight"] = _conv_default((ndf * 4, ndf * 2, 4, 4, 4), gen)
        _bn("main.6", ndf * 4, P, gen, False)
        P["main.9.weight"] = _conv_default((1, ndf * 4, 4, 4, 4), gen)
    else:
        chans = [C, ndf, ndf * 2, ndf * 4]
        for i in range(3):
            P[f"main.{4 * i + 1}.weight"] = _conv_default((chans[i + 1], chans[i], 4, 4, 4), gen)
            _bn(f"main.{4 * i + 2}", chans[i + 1], P, gen, False)
        P["main.13.weight"] = _conv_default((1, ndf * 4, 4, 4, 4), gen)
    return P


def is_param(key):
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked"))


def require_grad(P):
    for k, v in P.items():
        if is_param(k):
            v.requires_grad_(True)
    return P


# ------------------------------------------------------------------------------------------ layers
def _batch_norm(P, prefix, x, training):
    """nn.BatchNorm2d/3d forward incl. running-stat update and num_batches_tracked (torch semantics)."""
    if training:
        P[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"], P[prefix + ".weight"],
                        P[prefix + ".bias"], training, BN_MOM, BN_EPS)


def _noise(x, use_noise, sigma):
    """discriminator.py:30-39 - active regardless of train/eval; one normal_ draw of x's shape."""
    if use_noise:
        return x + sigma * torch.empty(x.size()).normal_()
    return x


# ------------------------------------------------------------------------------------------ ggen
def ggen_sample_z(P, B, cfg):
    """generator.py:84-116.  Draw order: z_content, h0, e_1..e_T (all torch.empty().normal_())."""
    g, T = cfg["ggen"], cfg["video_length"]
    z_c = torch.empty((B, g["dim_z_content"])).normal_()
    z_c = z_c.repeat(1, T).view(B * T, -1)                       # generator.py:103-108
    h = torch.empty((B, g["dim_z_motion"])).normal_()             # generator.py:84-85
    hs = []
    for _ in range(T):
        e = torch.empty((B, g["dim_z_motion"])).normal_()         # generator.py:87-88
        # nn.GRUCell.forward == torch.gru_cell; `gru_cell` below restates its equations (tests/test_oracle.py checks both agree)
        h = torch.gru_cell(e, h, P["recurrent.weight_ih"], P["recurrent.weight_hh"], P["recurrent.bias_ih"], P["recurrent.bias_hh"])
        hs.append(h)
    z_m = torch.stack(hs, 1).view(B * T, -1)                      # generator.py:96-99
    return torch.cat([z_c, z_m], 1)                               # generator.py:114


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """nn.GRUCell equations (gate order r, z, n)."""
    gi, gh = F.linear(x, w_ih, b_ih), F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def ggen_main(P, z, cfg, training):
    """generator.py:60-80: 5 x ConvTranspose2d, 4 x (BN + ReLU), Tanh / Softmax(dim=1)."""
    h = z.view(z.shape[0], -1, 1, 1)
    for i in range(5):
        stride, pad = (1, 0) if i == 0 else (2, 1)
        h = F.conv_transpose2d(h, P[f"main.{3 * i}.weight"], stride=stride, padding=pad)
        if i < 4:
            h = F.relu(_batch_norm(P, f"main.{3 * i + 1}", h, training))
    if cfg["geometric_info"]["name"] == "segmentation":
        return torch.softmax(h, 1)
    return torch.tanh(h)


def ggen_sample_videos(P, B, cfg, training=True):
    """generator.py:118-141 -> (B, C, T, 64, 64) view of (B, T, C, H, W) memory."""
    T, C = cfg["video_length"], cfg["geometric_info"]["channel"]
    h = ggen_main(P, ggen_sample_z(P, B, cfg), cfg, training)
    return h.view(B, T, C, 64, 64).permute(0, 2, 1, 3, 4)


# ------------------------------------------------------------------------------------------ cgen
def cgen_forward(P, x, z, cfg, training=True):
    """generator.py:361-402 (blocks :171-176,:203-207,:238-248,:272-277)."""
    if cfg["geometric_info"]["name"] == "segmentation":           # generator.py:378-385
        idx = torch.argmax(x, 1, keepdim=True)
        x = torch.full_like(x, -1.0).scatter_(1, idx, 1.0)
    hs = [F.leaky_relu(F.conv2d(x, P["inconv.main.0.weight"], padding=1), 0.01)]
    for i in range(6):
        h = F.conv2d(hs[-1], P[f"down_blocks.{i}.main.0.weight"], stride=2, padding=1)
        hs.append(F.leaky_relu(_batch_norm(P, f"down_blocks.{i}.main.1", h, training), 0.2))
    h = torch.cat([hs[-1], z], 1)                                 # generator.py:393
    for i in range(6):
        if i > 0:
            h = torch.cat([h, hs[-i - 1]], 1)                     # generator.py:398
        h = F.conv_transpose2d(h, P[f"up_blocks.{i}.main.0.weight"], stride=2, padding=1)
        h = _batch_norm(P, f"up_blocks.{i}.main.1", h, training)
        if i < 2:
            h = F.dropout2d(h, 0.5, training)                     # generator.py:246-248,337-338
        h = F.relu(h)
    h = F.conv_transpose2d(torch.cat([h, hs[0]], 1), P["outconv.main.0.weight"], stride=1, padding=1)
    return torch.tanh(h)                                          # generator.py:272-277,400


def cgen_forward_videos(P, xs, cfg, training=True):
    """generator.py:404-435: one z_color per clip, repeated over T; T folded into the batch."""
    B, C, T, H, W = xs.shape
    z = torch.empty((B, cfg["cgen"]["dim_z_color"])).normal_()    # generator.py:355-359
    zs = z.view(B, 1, -1, 1, 1).repeat(1, T, 1, 1, 1).view(B * T, -1, 1, 1)
    x = xs.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    ys = cgen_forward(P, x, zs, cfg, training)
    return ys.view(B, T, 3, H, W).permute(0, 2, 1, 3, 4)


# ------------------------------------------------------------------------------------------ discriminators
def idis_forward(P, xg, xc, cfg, training=True):
    """discriminator.py:79-102,106-127."""
    d = cfg["idis"]
    un, sg = d["use_noise"], d["noise_sigma"]
    hg = F.leaky_relu(F.conv2d(_noise(xg, un, sg), P["conv_g.1.weight"], stride=2, padding=1), 0.2)
    hc = F.leaky_relu(F.conv2d(_noise(xc, un, sg), P["conv_c.1.weight"], stride=2, padding=1), 0.2)
    h = torch.cat([hc, hg], 1)                                    # colour first, discriminator.py:124
    for conv, bn in (("main.1", "main.2"), ("main.5", "main.6")):
        h = F.conv2d(_noise(h, un, sg), P[conv + ".weight"], stride=2, padding=1)
        h = F.leaky_relu(_batch_norm(P, bn, h, training), 0.2)
    h = F.conv2d(_noise(h, un, sg), P["main.9.weight"], stride=2, padding=1)
    return h.squeeze()


_S3, _P3 = (1, 2, 2), (0, 1, 1)


def vdis_forward(P, xg, xc, cfg, training=True):
    """discriminator.py:180-207,210-231 (no Noise on the stems)."""
    d = cfg["vdis"]
    un, sg = d["use_noise"], d["noise_sigma"]
    hg = F.leaky_relu(F.conv3d(xg, P["conv_g.0.weight"], stride=_S3, padding=_P3), 0.2)
    hc = F.leaky_relu(F.conv3d(xc, P["conv_c.0.weight"], stride=_S3, padding=_P3), 0.2)
    h = torch.cat([hc, hg], 1)
    for conv, bn in (("main.1", "main.2"), ("main.5", "main.6")):
        h = F.conv3d(_noise(h, un, sg), P[conv + ".weight"], stride=_S3, padding=_P3)
        h = F.leaky_relu(_batch_norm(P, bn, h, training), 0.2)
    h = F.conv3d(_noise(h, un, sg), P["main.9.weight"], stride=_S3, padding=_P3)
    return h.squeeze()


def gdis_forward(P, xg, xc, cfg, training=True):
    """discriminator.py:285-306,309-333: temporal finite difference of xg only; xc unused."""
    d = cfg["gdis"]
    un, sg = d["use_noise"], d["noise_sigma"]
    L = xg.shape[2]
    h = xg[:, :, 1:L] - xg[:, :, 0:L - 1]
    for i in range(3):
        h = F.conv3d(_noise(h, un, sg), P[f"main.{4 * i + 1}.weight"], stride=_S3, padding=_P3)
        h = F.leaky_relu(_batch_norm(P, f"main.{4 * i + 2}", h, training), 0.2)
    h = F.conv3d(_noise(h, un, sg), P["main.13.weight"], stride=_S3, padding=_P3)
    return h.squeeze()


DIS_FORWARD = {"idis": idis_forward, "vdis": vdis_forward, "gdis": gdis_forward}


# ------------------------------------------------------------------------------------------ losses
def _bce_mean(y, target):
    return F.binary_cross_entropy_with_logits(y, torch.full_like(y, target), reduction="sum") / y.numel()


def dis_loss(kind, y_real, y_fake):
    """loss.py:93-99 (adversarial), :163-166 (hinge)."""
    if kind == "adversarial-loss":
        return _bce_mean(y_real, 1.0) + _bce_mean(y_fake, 0.0)
    return torch.mean(F.relu(1.0 - y_real)) + torch.mean(F.relu(1.0 + y_fake))


def gen_loss(kind, y_i, y_v, y_g):
    """loss.py:123-131 (adversarial: i + v + g), :190-193 (hinge: softplus(-y) of i and v only)."""
    if kind == "adversarial-loss":
        loss = _bce_mean(y_i, 1.0) + _bce_mean(y_v, 1.0)
        if y_g is not None:
            loss = loss + _bce_mean(y_g, 1.0)
        return loss
    return torch.mean(F.softplus(-y_i)) + torch.mean(F.softplus(-y_v))


# ------------------------------------------------------------------------------------------ training step
class OracleTrainer:
    """State + one iteration of trainer.py:271-363, with the optimizers of train.py:167-176.

    The step is executed *as written* by the reference: the D-phase backward runs through the
    non-detached fakes into the generators, the G-phase backward computes discriminator weight
    gradients, opt_ggen.step() is called twice, the update gates use the swapped config names.
    """

    NETS = ("ggen", "cgen", "idis", "vdis", "gdis")

    def __init__(self, cfg, params):
        self.cfg = cfg
        self.P = {k: require_grad(v) for k, v in params.items()}
        self.use_gdis = cfg["gdis"].get("enabled", True)
        self.opt = {}
        for name in self.NETS:
            if name == "gdis" and not self.use_gdis:
                continue
            o = cfg[name]["optimizer"]
            plist = [v for k, v in self.P[name].items() if is_param(k)]
            self.opt[name] = torch.optim.Adam(plist, lr=o["lr"], betas=(0.5, 0.999), weight_decay=o["decay"])
        self.iteration = 0
        self.capture_grads = False
        self.d_grads = self.g_grads = None
        self.gen_training = True   # log_samples()/evaluate() leave the generators in eval mode (trainer.py:126-127)

    def _zero(self, names):
        for n in names:
            if n in self.opt:
                self.opt[n].zero_grad()

    def _dis_all(self, xg, xc, t_rand):
        cfg, P = self.cfg, self.P
        y_i = idis_forward(P["idis"], xg[:, :, t_rand], xc[:, :, t_rand], cfg)   # trainer.py:299,307,347
        y_v = vdis_forward(P["vdis"], xg, xc, cfg)
        y_g = gdis_forward(P["gdis"], xg, xc, cfg) if self.use_gdis else None
        return y_i, y_v, y_g

    def step(self, xc_real, xg_real, t_rand=None):
        cfg, P = self.cfg, self.P
        B, kind = cfg["batchsize"], cfg["loss"]
        self.iteration += 1
        if t_rand is None:
            t_rand = np.random.randint(cfg["video_length"])                       # trainer.py:279
        # ---- discriminator phase (trainer.py:285-333)
        self._zero(("idis", "vdis", "gdis"))
        yr_i, yr_v, yr_g = self._dis_all(xg_real, xc_real, t_rand)
        xg_fake = ggen_sample_videos(P["ggen"], B, cfg, self.gen_training)        # not detached, trainer.py:304
        xc_fake = cgen_forward_videos(P["cgen"], xg_fake, cfg, self.gen_training)
        yf_i, yf_v, yf_g = self._dis_all(xg_fake, xc_fake, t_rand)
        loss_i, loss_v = dis_loss(kind, yr_i, yf_i), dis_loss(kind, yr_v, yf_v)
        loss_dis = loss_i + loss_v
        loss_g = None
        if self.use_gdis:
            loss_g = dis_loss(kind, yr_g, yf_g)
            loss_dis = loss_dis + loss_g
        if self.iteration % cfg["num_gen_update"] == 0:                           # trainer.py:318 (names swapped)
            loss_dis.backward()
            if self.capture_grads:   # test hook: the useful D gradients, before the G-phase adds its dead ones
                self.d_grads = {n: {k: v.grad.clone() for k, v in self.P[n].items() if is_param(k)}
                                for n in ("idis", "vdis", "gdis") if n in self.opt}
            for n in ("idis", "vdis", "gdis"):
                if n in self.opt:
                    self.opt[n].step()
        out = {"loss_idis": float(loss_i.detach()), "loss_vdis": float(loss_v.detach()),
               "loss_gdis": float(loss_g.detach()) if loss_g is not None else None}
        # ---- generator phase (trainer.py:338-368)
        self.gen_training = True
        self._zero(("ggen", "cgen"))
        xg_fake = ggen_sample_videos(P["ggen"], B, cfg, True)
        xc_fake = cgen_forward_videos(P["cgen"], xg_fake, cfg, True)
        yf_i, yf_v, yf_g = self._dis_all(xg_fake, xc_fake, t_rand)
        loss_gen = gen_loss(kind, yf_i, yf_v, yf_g)
        if self.iteration % cfg["num_dis_update"] == 0:                           # trainer.py:355
            loss_gen.backward()
            if self.capture_grads:
                self.g_grads = {n: {k: v.grad.clone() for k, v in self.P[n].items() if is_param(k)} for n in ("ggen", "cgen")}
            self.opt["ggen"].step()
            self.opt["cgen"].step()
            self.opt["ggen"].step()                                               # trainer.py:357-359 (twice)
        out["loss_gen"] = float(loss_gen.detach())
        out["t_rand"] = int(t_rand)
        return out


def synthetic_batch(cfg, B, seed):
    """SURVEY.md section 8d synthetic inputs: colour U(-1,1); geometry U(-1,1) or one-hot(randint)."""
    gen = torch.Generator().manual_seed(seed)
    T, C = cfg["video_length"], cfg["geometric_info"]["channel"]
    xc = torch.rand((B, 3, T, 64, 64), generator=gen) * 2 - 1
    if cfg["geometric_info"]["name"] == "segmentation":
        idx = torch.randint(0, C, (B, 1, T, 64, 64), generator=gen)
        xg = torch.zeros((B, C, T, 64, 64)).scatter_(1, idx, 1.0)
    else:
        xg = torch.rand((B, C, T, 64, 64), generator=gen) * 2 - 1
    return xc, xg


def init_all(cfg, seed):
    gen = torch.Generator().manual_seed(seed)
    P = {"ggen": init_ggen(cfg, gen), "cgen": init_cgen(cfg, gen)}
    for k in ("idis", "vdis", "gdis"):
        P[k] = init_dis(k, cfg, gen)
    return P
