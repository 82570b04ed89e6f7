"""BASELINE.json configs[1] and [2] with the model IN the loop (random-init weights, synthetic probes; no network):

  c2  EfficientNet-B5 (the reference's own class, staged under baseline/_ref by tools/stage_reference.py), all 39 MBConv
      blocks hooked, synthetic 1520 x 912 probes                   -> M-Mammo-CLIP Dissect shape
  c3  HF CLIP ViT-B/16 vision tower, the 12 encoder layers hooked, synthetic 224 x 224 probes   -> G-Mammo-CLIP shape

The forward hooks are this repo's (hooks.ActivationStack: K4 pools every hooked NCHW activation straight into one
device-resident [N, sum K_l] matrix; [B, T, D] layers yield the CLS token), the encoder forward is stock PyTorch.
Printed separately, as the north star asks: encoder forward time (hooks removed), time inside the hooks (K4, CUDA events
around every hook call) with the bytes it read, and the scoring time (soft_wpmi_layers over all layers in one pass, and
layer by layer), plus -- as a cross-check of the pooled matrix -- the max deviation from torch's own mean.

    python tools/run_configs.py c2 [--probes 5000] [--batch 4] [--json out.json]
    python tools/run_configs.py c3 [--probes 10000] [--batch 250]
"""
import argparse
import importlib
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mammo_clip_dissect_b200 import features, hooks, similarity  # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref")


def reference_efficientnet_b5():
    """The reference's EfficientNet class (model/modules/efficientnet_custom.py:143-313), imported from the staged copy
    through stand-in parent packages (model/__init__.py pulls in dependencies that are not installed here)."""
    mod_dir = os.path.join(REF, "model", "modules")
    if not os.path.exists(os.path.join(mod_dir, "efficientnet_custom.py")):
        raise SystemExit("reference not staged: run python tools/stage_reference.py")
    for name, path in (("model", os.path.join(REF, "model")), ("model.modules", mod_dir)):
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    eff = importlib.import_module("model.modules.efficientnet_custom")
    torch.manual_seed(0)
    return eff.EfficientNet.from_name("efficientnet-b5").eval()


def clip_vit_b16():
    from transformers import CLIPVisionConfig, CLIPVisionModel
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                           image_size=224, patch_size=16)
    return CLIPVisionModel(cfg).eval()


def _unwrap(o):
    return o[0] if type(o) is tuple else o


class TimedHook:
    """Wraps a forward hook with a pair of CUDA events and counts the bytes of the activation it pooled."""

    def __init__(self, fn):
        self.fn, self.events, self.bytes = fn, [], 0

    def __call__(self, module, inp, out):
        x = _unwrap(out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.fn(module, inp, out)
        e1.record()
        self.events.append((e0, e1))
        if x.dim() == 4:
            self.bytes += x.numel() * x.element_size()

    def ms(self):
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.events)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c2", "c3"])
    ap.add_argument("--probes", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--json", default=None)
    ap.add_argument("--check-batches", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.config == "c2":
        model, layers = reference_efficientnet_b5(), None
        n, batch, shape, seed = args.probes or 5000, args.batch or 4, (3, 1520, 912), 4
        model = model.to(dev)
        layers = list(model._blocks)
    else:
        model = clip_vit_b16().to(dev)
        layers = list(model.vision_model.encoder.layers)
        n, batch, shape, seed = args.probes or 10000, args.batch or 250, (3, 224, 224), 5
    n = n // batch * batch
    g = torch.Generator(device=dev).manual_seed(seed)
    # widths of the hooked layers from one dry forward
    widths, handles = [], []
    for l in layers:
        handles.append(l.register_forward_hook(
            lambda m, i, o: widths.append(_unwrap(o).shape[1 if _unwrap(o).dim() == 4 else 2])))
    with torch.no_grad():
        model(torch.randn((1,) + shape, generator=g, device=dev))
    for h in handles:
        h.remove()
    print("%s: %d hooked layers, widths %s (sum %d), %d probes of %s in batches of %d"
          % (args.config, len(layers), sorted(set(widths)), sum(widths), n, "x".join(map(str, shape)), batch), flush=True)

    # ---- encoder forward alone (stock PyTorch, no hooks) ------------------------------------------------------------
    x = torch.randn((batch,) + shape, generator=g, device=dev)
    with torch.no_grad():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            model(x)
        e1.record()
        torch.cuda.synchronize()
    fwd_ms_batch = e0.elapsed_time(e1) / reps

    # ---- the probe pass with the hooks in place: K4 writes into the stacked matrix ------------------------------------
    stack = hooks.ActivationStack(n, widths, dev)
    timed = [TimedHook(stack.hook(i, "avg")) for i in range(len(layers))]
    handles = [l.register_forward_hook(t) for l, t in zip(layers, timed)]
    check = []
    if args.check_batches:
        handles += [l.register_forward_hook(lambda m, i, o, k=k: check.append(
            (k, _unwrap(o).float().mean(dim=[2, 3]) if _unwrap(o).dim() == 4 else _unwrap(o)[:, 0].float())))
            for k, l in enumerate(layers)]
    t0 = time.perf_counter()
    max_dev = 0.0
    with torch.no_grad():
        for b in range(n // batch):
            x = torch.randn((batch,) + shape, generator=g, device=dev)
            model(x)
            if check:
                for k, ref in check:
                    got = stack.layer(k)[b * batch:(b + 1) * batch]
                    max_dev = max(max_dev, ((got - ref).abs().max() / ref.abs().mean().clamp_min(1e-30)).item())
                check.clear()
                if b + 1 >= args.check_batches:
                    for h in handles[len(layers):]:
                        h.remove()
                    handles = handles[:len(layers)]
    torch.cuda.synchronize()
    wall_s = time.perf_counter() - t0
    hook_ms = sum(t.ms() for t in timed)
    hook_bytes = sum(t.bytes for t in timed)
    for h in handles:
        h.remove()
    assert stack.complete()

    # ---- scoring: clip_feats from synthetic image / text features (K1), then soft-WPMI over all layers ----------------
    C = 763
    I = torch.randn(n, 512, generator=g, device=dev)
    T = torch.randn(C, 512, generator=g, device=dev)

    def timeit(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters, out

    k1_ms, P = timeit(lambda: features.similarity_matrix(I, T, device=dev))
    one_ms, stacked = timeit(lambda: similarity.soft_wpmi_layers(P, stack, device=dev))
    per_ms, separate = timeit(lambda: [similarity.soft_wpmi(P, stack.layer(i), device=dev) for i in range(len(layers))])
    same = all(torch.equal(a, b) for a, b in zip(stacked, separate))
    top_ms, _ = timeit(lambda: [similarity.top_concepts(s, 10) for s in stacked])
    res = {"config": args.config, "probes": n, "batch": batch, "layers": len(layers), "neurons": int(sum(widths)),
           "encoder_forward_ms_per_batch": round(fwd_ms_batch, 3),
           "encoder_forward_s_total": round(fwd_ms_batch * (n // batch) / 1e3, 3),
           "probe_pass_wall_s_with_hooks": round(wall_s, 3),
           "k4_hook_ms_total": round(hook_ms, 3), "k4_bytes_read": int(hook_bytes),
           "k4_gbs": round(hook_bytes / max(hook_ms, 1e-9) / 1e6, 1) if hook_bytes else None,
           "k4_calls": len(layers) * (n // batch),
           "k4_vs_torch_mean_max_rel_dev": max_dev,
           "k1_similarity_matrix_ms": round(k1_ms, 4),
           "soft_wpmi_all_layers_one_pass_ms": round(one_ms, 4), "soft_wpmi_layer_by_layer_ms": round(per_ms, 4),
           "one_pass_equals_layer_by_layer_bitwise": bool(same), "top10_concepts_ms": round(top_ms, 4),
           "neurons_per_s_scoring": round(sum(widths) / (one_ms / 1e3), 1)}
    print(json.dumps(res), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
