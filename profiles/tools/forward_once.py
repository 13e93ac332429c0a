"""One eager UNet window-forward (BASELINE configs[1] shape: b = 2, f = 16, 64x64 latents, full width) and one VAE decode
of 16 frames inside a cudaProfilerStart/Stop region, plus a sidecar JSON that says, for every kernel this repo launched in
that region IN LAUNCH ORDER, which operator it belongs to, its shapes and its ALGORITHMIC FLOPs and bytes.

  ncu --profile-from-start off --metrics <...> --csv --log-file gpurun_out/r02_roofline_raw.csv \
      python profiles/tools/forward_once.py gpurun_out/r02_oplog.json
  python profiles/tools/roofline_merge.py gpurun_out/r02_roofline_raw.csv gpurun_out/r02_oplog.json profiles/r02_roofline.csv
"""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from vexpress_b200 import _ffi, ops  # noqa: E402

LOG, CUR = [], [None]


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def meta_of(name, a, k):
    """(shape tag, algorithmic flop, algorithmic bytes) -- bytes = every operand and result once."""
    if name == "gemm":
        A, W = a[0], a[1]
        a2, res = k.get("a2"), k.get("residual")
        M, K, N = A.shape[0], A.shape[1] + (a2.shape[1] if a2 is not None else 0), W.shape[0]
        nout = N // 2 if k.get("geglu") else N
        ob = 4 if k.get("out_f32") else 2
        return (f"M={M} K={K} N={N}" + (" geglu" if k.get("geglu") else "") + (" +res" if res is not None else ""),
                2.0 * M * N * K, 2.0 * (M * K + N * K) + ob * M * nout + (2.0 * M * nout if res is not None else 0))
    if name in ("gemm_rowsums", "gemm_lnparts"):   # LayerNorm statistics hand-over: producer (+ 8 B x slots per row) / consumer
        A, W = a[0], a[1]
        res = k.get("residual")
        M, K, N = A.shape[0], A.shape[1], W.shape[0]
        nout = N // 2 if k.get("geglu") else N
        tag = "rowsums" if name == "gemm_rowsums" else "lnparts"
        return (f"M={M} K={K} N={N} {tag}" + (" geglu" if k.get("geglu") else "") + (" +res" if res is not None else ""),
                2.0 * M * N * K, 2.0 * (M * K + N * K) + 2.0 * M * nout + (2.0 * M * nout if res is not None else 0))
    if name == "conv3x3":
        X, W = a[0], a[1]
        nb, h, w, c = X.shape
        co = W.shape[0]
        M = nb * h * w
        res = k.get("residual")
        return (f"{nb}x{h}x{w} C={c}->{co}" + (" +res" if res is not None else ""), 2.0 * M * co * 9 * c,
                2.0 * (M * c + 9 * c * co + M * co) + (2.0 * M * co if res is not None else 0))
    if name == "upconv3x3":   # conv3x3(upsample2x(x)): canonical 9-tap FLOPs on the 2x image (the kernel executes 4/9 of them)
        X, W4 = a[0], a[1]
        nb, h, w, c = X.shape
        co = W4.shape[0] // 4
        M = nb * h * w
        return (f"{nb}x{h}x{w}->2x C={c}->{co}", 2.0 * (4 * M) * co * 9 * c, 2.0 * (M * c + 9 * c * co + 4 * M * co))
    if name == "conv3x3_s2":
        X, W = a[0], a[1]
        nb, h, w, c = X.shape
        co = W.shape[0]
        M = nb * (h // 2) * (w // 2)
        return (f"{nb}x{h}x{w} C={c}->{co} stride 2", 2.0 * M * co * 9 * c, 2.0 * (nb * h * w * c + 9 * c * co + M * co))
    if name == "flash_attention":
        q, kk, heads, Nq, Nk = a[0], a[1], a[3], a[4], a[5]
        C = q.shape[1]
        return (f"Bq={q.shape[0] // Nq} Nq={Nq} Nk={Nk} hd={C // heads} kv_div={k.get('kv_div', a[6] if len(a) > 6 else 1)}",
                4.0 * q.shape[0] * Nk * C, 2.0 * (2 * q.shape[0] * C + 2 * kk.shape[0] * C))
    if name == "groupnorm":
        x1, x2 = a[0], k.get("x2")
        C = x1.shape[1] + (x2.shape[1] if x2 is not None else 0)
        return (f"rows={x1.shape[0]} C={C}" + (" silu" if a[6] else ""), 0.0, None)      # per-kernel bytes filled in check()
    if name == "layernorm":
        x = a[0]
        return (f"rows={x.shape[0]} C={x.shape[1]}" + (" +pe" if k.get("pe") is not None else ""), 0.0, 4.0 * x.numel())
    if name == "temporal_attention":
        q = a[0]
        f = a[4]
        return (f"rows={q.shape[0]} C={q.shape[1]} f={f}", 4.0 * q.shape[0] * f * q.shape[1], 8.0 * q.numel())
    if name == "smallkv_attention":
        q = a[0]
        return (f"rows={q.shape[0]} C={q.shape[1]} Lk={a[5]}", 4.0 * q.shape[0] * a[5] * q.shape[1], 4.0 * q.numel())
    if name in ("upsample2x", "im2col_s2", "im2col3x3"):
        x = a[0]
        mult = 4 if name == "upsample2x" else 9 / 4
        return (f"rows={x.shape[0] if x.dim() == 2 else x.numel() // x.shape[-1]} C={x.shape[-1]}", 0.0, 2.0 * x.numel() * (1 + mult))
    if name == "conv_in":
        x = a[0]
        return (f"{tuple(x.shape)} -> {a[3]}", 2.0 * x.shape[0] * x.shape[2] * x.shape[3] * 36 * a[3],
                2.0 * x.numel() + 2.0 * x.shape[0] * x.shape[2] * x.shape[3] * a[3] * (2 if k.get("addend") is not None else 1))
    if name == "softmax_rows":
        x = a[0]
        return (f"{tuple(x.shape)}", 0.0, 6.0 * x.numel())
    return ("", 0.0, 0.0)


NAMES = ["gemm", "gemm_rowsums", "gemm_lnparts", "conv3x3", "conv3x3_s2", "upconv3x3", "flash_attention", "temporal_attention", "smallkv_attention", "groupnorm", "layernorm", "conv_in",
         "conv_out_tc", "im2col_s2", "im2col3x3", "upsample2x", "skinny_linear", "timestep_embed", "softmax_rows", "geglu",
         "cfg_overlap_accumulate", "ddim_step"]


def install():
    real_check = _ffi.check

    def check(rc, what=""):
        real_check(rc, what)
        cur = CUR[0]
        if cur is None:
            return
        flop, nbytes = cur["flop"], cur["bytes"]
        if cur["op"] == "groupnorm":      # statistics kernel: one read; apply kernel: one read + one write
            nbytes = cur["gn"] * (1.0 if what.endswith("stats") else 2.0)
        LOG.append(dict(op=cur["op"], entry=what, shape=cur["shape"], flop=flop, bytes=nbytes))
        cur["flop"], cur["bytes"] = 0.0, 0.0          # further launches of the same operator call carry no extra work
    _ffi.check = check
    ops.check = check
    for n in NAMES:
        def mk(n, fn):
            def w(*a, **k):
                outer = CUR[0]
                if outer is None:
                    tag, flop, nb = meta_of(n, a, k)
                    gn = 0.0
                    if n == "groupnorm":
                        x2 = k.get("x2")
                        gn = 2.0 * (a[0].numel() + (x2.numel() if x2 is not None else 0))
                    CUR[0] = dict(op=n, shape=tag, flop=flop, bytes=nb or 0.0, gn=gn)
                try:
                    return fn(*a, **k)
                finally:
                    if outer is None:
                        CUR[0] = None
            return w
        setattr(ops, n, mk(n, getattr(ops, n)))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_oplog.json")
    L, h = 16, 64
    pipe, host, banks = bench.build_ours(L, h, torch.device("cuda", 0))
    from vexpress_b200.modules import ReferenceAttentionControl
    reader = ReferenceAttentionControl(pipe.denoising_unet, do_classifier_free_guidance=True, mode="read", fusion_blocks="full",
                                       reference_attention_weight=0.95, audio_attention_weight=3.0)
    reader.update(pipe.reference_net.writer_view, True, dtype=torch.bfloat16)
    eng = pipe.denoising_unet.engine()
    frames = host["lat"].cuda()[0].permute(1, 0, 2, 3).repeat(2, 1, 1, 1).contiguous()
    enc = host["audio"].cuda().reshape(2 * L, 5, 768)
    kps = host["kps"].cuda().permute(0, 2, 3, 4, 1).reshape(2 * L * h * h, 320).contiguous()
    z = host["lat"].cuda()[0].permute(1, 0, 2, 3).contiguous()
    for _ in range(2):
        eng.forward_frames(frames, 499, enc, kps, None, 2, L)
        pipe.vae.decode_latents(z)
    torch.cuda.synchronize()
    install()
    torch.cuda.profiler.start()
    eng.forward_frames(frames, 499, enc, kps, None, 2, L)
    n_unet = len(LOG)
    pipe.vae.decode_latents(z)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    for i, r in enumerate(LOG):
        r["part"] = "unet" if i < n_unet else "vae"
    json.dump(LOG, open(out, "w"))
    print(f"{len(LOG)} launches logged ({n_unet} UNet, {len(LOG) - n_unet} VAE) -> {out}")


if __name__ == "__main__":
    main()
