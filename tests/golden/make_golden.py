"""Generates the golden vectors that pin the CPU oracle (and the CUDA kernels) to the REFERENCE.

Part 1 (needs a GPU, run on the B200 box: `python tests/golden/make_golden.py --gpu`):
    runs the reference's own CUDA extension (baseline/_ref, built by oracle/build_ref.sh from
    /root/reference with the 4-line torch-2 compat patch) through its public primitives
    rubiks_shift_3d_forward/backward and rubiks2d_forward/backward on small seeded inputs and stores
    inputs + outputs in gpurun_out/golden/shift_golden.npz (then copied to tests/golden/).
Part 2 (CPU, run in the build container: `python tests/golden/make_golden.py --cpu`):
    * AttentionShift: imports the reference's pure-PyTorch module from /root/reference and records
      forward output and autograd gradients -> attention_golden.npz
    * the 51 trained shift tensors of pretrained/ssv2_large.pth.tar -> pretrained_shifts_large.npz
      (realistic shift distributions incl. exact-integer entries).
Nothing here is imported by the product package.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def cases3d():
    """name -> dict(shape, stride, padding, shift-kind, quantize, normalize_grad, t_factor, dtype)"""
    c = {}
    c["s1_rand"] = dict(shape=(2, 4, 6, 9, 10), stride=(1, 1, 1), padding=(0, 0, 0), kind="rand1.5")
    c["s1_rand_nonorm"] = dict(shape=(2, 4, 6, 9, 10), stride=(1, 1, 1), padding=(0, 0, 0), kind="rand1.5",
                               normalize_grad=False)
    c["s122_rand"] = dict(shape=(2, 4, 5, 12, 11), stride=(1, 2, 2), padding=(0, 0, 0), kind="rand1.5",
                          normalize_grad=False)
    c["s221_p110"] = dict(shape=(2, 5, 4, 8, 9), stride=(2, 2, 1), padding=(1, 1, 0), kind="rand1.5",
                          normalize_grad=False)
    c["s1_big_shift"] = dict(shape=(1, 8, 8, 7, 7), stride=(1, 1, 1), padding=(0, 0, 0), kind="rand6",
                             normalize_grad=False)
    c["s1_integer"] = dict(shape=(2, 4, 8, 6, 7), stride=(1, 1, 1), padding=(0, 0, 0), kind="integer",
                           normalize_grad=False)
    c["s122_integer"] = dict(shape=(2, 4, 8, 8, 10), stride=(1, 2, 2), padding=(0, 0, 0), kind="integer",
                             normalize_grad=False)
    c["s1_half"] = dict(shape=(2, 3, 6, 6, 6), stride=(1, 1, 1), padding=(0, 0, 0), kind="halves",
                        normalize_grad=False)
    c["s1_quant"] = dict(shape=(2, 4, 8, 6, 7), stride=(1, 1, 1), padding=(0, 0, 0), kind="halves", quantize=True)
    c["s122_quant"] = dict(shape=(2, 4, 6, 8, 10), stride=(1, 2, 2), padding=(0, 0, 0), kind="rand1.5",
                           quantize=True)
    c["s1_tfactor_neg"] = dict(shape=(2, 4, 6, 9, 10), stride=(1, 1, 1), padding=(0, 0, 0), kind="rand1.5",
                               t_factor=-1.0)
    c["s1_tfactor_half"] = dict(shape=(2, 4, 6, 9, 10), stride=(1, 1, 1), padding=(0, 0, 0), kind="rand1.5",
                                t_factor=0.5)
    c["s1_double"] = dict(shape=(2, 4, 6, 9, 10), stride=(1, 1, 1), padding=(0, 0, 0), kind="rand1.5",
                          dtype="float64", normalize_grad=False)
    c["s122_double_int"] = dict(shape=(1, 4, 8, 8, 10), stride=(1, 2, 2), padding=(0, 0, 0), kind="integer",
                                dtype="float64", normalize_grad=False)
    return c


def cases2d():
    c = {}
    c["s1_rand"] = dict(shape=(3, 6, 9, 10), stride=(1, 1), padding=(0, 0), kind="rand1.5")
    c["s2_rand"] = dict(shape=(3, 5, 12, 11), stride=(2, 2), padding=(0, 0), kind="rand1.5", normalize_grad=False)
    c["s21_p10"] = dict(shape=(2, 4, 8, 9), stride=(2, 1), padding=(1, 0), kind="rand1.5", normalize_grad=False)
    c["s1_integer"] = dict(shape=(2, 8, 6, 7), stride=(1, 1), padding=(0, 0), kind="integer", normalize_grad=False)
    c["s2_integer"] = dict(shape=(2, 8, 8, 10), stride=(2, 2), padding=(0, 0), kind="integer", normalize_grad=False)
    c["s1_quant"] = dict(shape=(2, 8, 6, 7), stride=(1, 1), padding=(0, 0), kind="halves", quantize=True)
    c["s2_quant"] = dict(shape=(2, 6, 8, 10), stride=(2, 2), padding=(0, 0), kind="rand1.5", quantize=True)
    c["s1_double"] = dict(shape=(2, 6, 9, 10), stride=(1, 1), padding=(0, 0), kind="rand1.5", dtype="float64",
                          normalize_grad=False)
    c["s1_noshiftgrad"] = dict(shape=(2, 6, 9, 10), stride=(1, 1), padding=(0, 0), kind="rand1.5",
                               enable_shift_grad=False)
    return c


def make_shift(rng, kind, dims, C):
    if kind == "rand1.5":
        return rng.uniform(-1.5, 1.5, size=(dims, C))
    if kind == "rand6":
        return rng.uniform(-6, 6, size=(dims, C))
    if kind == "integer":  # every channel has >= 1 exactly-integer component, mixed with fractions
        vals = np.array([-2.0, -1.0, 0.0, 1.0, 2.0, 0.25, -0.75, 1.5])
        s = vals[rng.integers(0, len(vals), size=(dims, C))]
        s[0, :] = np.array([0.0, 1.0, -1.0, 0.3, 0.0, 2.0, -0.6, 1.0])[:C] if C <= 8 else s[0, :]
        s[:, 0] = 0.0  # an all-zero channel (zero-shift fast path of the input gradient)
        return s
    if kind == "halves":  # exactly .5 remainders: forward/backward quantize rounding differs there
        vals = np.array([-1.5, -0.5, 0.5, 1.5, 0.25, -0.25, 1.0, 0.0])
        return vals[rng.integers(0, len(vals), size=(dims, C))]
    raise ValueError(kind)


def out_len(n, s, p):
    return (n + 2 * p - 1) // s + 1


def gen_inputs(name, cfg, dims):
    rng = np.random.default_rng(sum(ord(ch) for ch in name) * 7919 + dims)  # stable across processes
    dtype = np.dtype(cfg.get("dtype", "float32"))
    shape = cfg["shape"]
    C = shape[2] if dims == 3 else shape[1]
    x = rng.standard_normal(shape).astype(dtype)
    shift = make_shift(rng, cfg["kind"], dims, C).astype(dtype)
    if dims == 3:
        oshape = (shape[0], out_len(shape[1], cfg["stride"][0], cfg["padding"][0]), C,
                  out_len(shape[3], cfg["stride"][1], cfg["padding"][1]),
                  out_len(shape[4], cfg["stride"][2], cfg["padding"][2]))
    else:
        oshape = (shape[0], C, out_len(shape[2], cfg["stride"][0], cfg["padding"][0]),
                  out_len(shape[3], cfg["stride"][1], cfg["padding"][1]))
    og = rng.standard_normal(oshape).astype(dtype)
    return x, shift, og


def run_gpu(out_path):
    import torch
    sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
    from rubiksnet.shiftlib.rubiks2d.primitive import rubiks2d_backward, rubiks2d_forward  # reference
    from rubiksnet.shiftlib.rubiks3d.primitive import rubiks_shift_3d_backward, rubiks_shift_3d_forward

    store = {}
    for name, cfg in cases3d().items():
        x, shift, og = gen_inputs(name, cfg, 3)
        tx, ts, tg = (torch.from_numpy(a).cuda() for a in (x, shift, og))
        q = cfg.get("quantize", False)
        out = rubiks_shift_3d_forward(tx, ts, list(cfg["stride"]), list(cfg["padding"]), quantize=q)
        assert tuple(out.shape) == og.shape, (name, out.shape, og.shape)
        gin, gs = rubiks_shift_3d_backward(tg, tx, ts, list(cfg["stride"]), list(cfg["padding"]),
                                           cfg.get("normalize_grad", True), cfg.get("t_factor", 1.0), q)
        torch.cuda.synchronize()
        for k, v in (("x", x), ("shift", shift), ("og", og), ("out", out.detach().cpu().numpy()),
                     ("gin", gin.detach().cpu().numpy()), ("gshift", gs.detach().cpu().numpy())):
            store["3d/%s/%s" % (name, k)] = v
    for name, cfg in cases2d().items():
        x, shift, og = gen_inputs(name, cfg, 2)
        tx, ts, tg = (torch.from_numpy(a).cuda() for a in (x, shift, og))
        q = cfg.get("quantize", False)
        out = rubiks2d_forward(tx, ts, list(cfg["stride"]), list(cfg["padding"]), quantize=q)
        gin, gs = rubiks2d_backward(tg, tx, ts, list(cfg["stride"]), list(cfg["padding"]),
                                    cfg.get("normalize_grad", True), cfg.get("enable_shift_grad", True), q)
        torch.cuda.synchronize()
        for k, v in (("x", x), ("shift", shift), ("og", og), ("out", out.detach().cpu().numpy()),
                     ("gin", gin.detach().cpu().numpy()), ("gshift", gs.detach().cpu().numpy())):
            store["2d/%s/%s" % (name, k)] = v
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **store)
    print("wrote", out_path, "with", len(store), "arrays from", torch.cuda.get_device_name(0))


def run_cpu():
    import torch
    ref_root = os.environ.get("RUBIKS_REFERENCE", "/root/reference")
    sys.path.insert(0, ref_root)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_attention_shift",
                                                  os.path.join(ref_root, "rubiksnet", "attention_shift.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    store = {}
    for name, (n, t, c, h, w, temp) in {"small": (2, 4, 6, 5, 7, 2.0), "sharp": (1, 8, 5, 3, 3, 0.05),
                                        "t1": (3, 1, 4, 2, 2, 2.0)}.items():
        torch.manual_seed(len(name) * 101 + c)
        m = mod.AttentionShift(t).double()
        x = torch.randn(n * t, c, h, w, dtype=torch.float64, requires_grad=True)
        m.weight = torch.nn.Parameter(torch.rand(c, 3, dtype=torch.float64))
        m.T.data.fill_(temp)
        out = m(x)
        og = torch.randn_like(out)
        out.backward(og)
        for k, v in (("x", x), ("weight", m.weight), ("og", og), ("out", out), ("gx", x.grad),
                     ("gweight", m.weight.grad)):
            store["%s/%s" % (name, k)] = v.detach().numpy()
        store["%s/meta" % name] = np.array([n, t, temp], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "attention_golden.npz"), **store)
    print("wrote attention_golden.npz")

    ck = torch.load(os.path.join(ref_root, "pretrained", "ssv2_large.pth.tar"), map_location="cpu")
    shifts = {k.replace(".", "/"): v.numpy() for k, v in ck["model"].items() if k.endswith("rubiks3d.shift")}
    np.savez_compressed(os.path.join(HERE, "pretrained_shifts_large.npz"), **shifts)
    print("wrote pretrained_shifts_large.npz with", len(shifts), "tensors")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "golden", "shift_golden.npz"))
    args = ap.parse_args()
    if args.gpu:
        run_gpu(args.out)
    if args.cpu:
        run_cpu()
