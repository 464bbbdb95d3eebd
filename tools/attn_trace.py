"""Per-step timeline of attention2_kernel on one SM (tools only).

CTA 0 records (clock64, tag, step) events for three roles into a device buffer handed over through MJV_ATTN_TRACE:
  role 0/1  softmax warp 0 of query tile 0/1: 0 before the s_full wait, 1 after it, 2 max known (before the exponentials),
            3 exponentials + P stores issued, 4 P stores complete (just before the p_full arrive)
  role 2    P·V issuer: 10 after the p_full wait of a step, 11 after issuing its P·V;  role 3  S issuer: 12 after issuing S of a step
Prints median cycle counts per phase. Usage: python tools/attn_trace.py [vit|llm]
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mj_video_b200 import ops  # noqa: E402

EV = 2048
dev = "cuda"


def run_vit(T=256, S=1025, H=16):
    Cv = H * 64
    qk = torch.randn(T * S, 2 * Cv, device=dev).bfloat16()
    vt = torch.randn(T * H, 64, (S + 7) // 8 * 8, device=dev).bfloat16()
    out = torch.empty(T * S, Cv, device=dev, dtype=torch.bfloat16)
    return lambda: ops.attention(qk, qk, vt, out, head_dim=64, heads=H, kv_heads=H, causal=False, scale=0.125,
                                 q_col0=0, k_col0=Cv, nseq=T, seq_len=S)


def run_llm(lens=None, H=16, Hkv=8, hd=128):
    lens = lens or [2187] * 32
    tot = sum(lens)
    q = torch.randn(tot, H * hd, device=dev).bfloat16()
    k = torch.randn(tot, Hkv * hd, device=dev).bfloat16()
    starts, cols = ops.vt_layout(lens)
    vt = torch.randn(Hkv, hd, cols, device=dev).bfloat16()
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=dev)
    vts = torch.tensor(starts, dtype=torch.int32, device=dev)
    work = ops.causal_work_list(lens).to(dev)
    out = torch.empty(tot, H * hd, device=dev, dtype=torch.bfloat16)
    return lambda: ops.attention(q, k, vt, out, head_dim=hd, heads=H, kv_heads=Hkv, causal=True, scale=1 / math.sqrt(hd),
                                 cu_seqlens=cu, work=work, vt_starts=vts, vt_cols=cols)


def summarize(buf, name):
    ev = buf.view(4, EV, 2).cpu().numpy().copy()
    res = {}
    for role in (0, 1):
        e = ev[role]
        n = int((e[:, 0] != 0).sum())
        clk, tag, step = e[:n, 0], e[:n, 1] >> 32, e[:n, 1] & 0xFFFFFFFF
        per = {}
        for s in np.unique(step):
            m = {int(t): int(c) for t, c in zip(tag[step == s], clk[step == s])}
            per[int(s)] = m
        steps = sorted(per)[8:]  # skip the pipeline fill
        def med(f):
            v = [f(per[s]) for s in steps if all(k in per[s] for k in (0, 1, 2, 3, 4))]
            return float(np.median(v)) if v else float("nan")
        period = np.median(np.diff([per[s][0] for s in steps if 0 in per[s]]))
        res[f"softmax{role}"] = {
            "events": n, "period": float(period), "wait_s": med(lambda m: m[1] - m[0]), "ld_max": med(lambda m: m[2] - m[1]),
            "exp": med(lambda m: m[3] - m[2]), "st_wait": med(lambda m: m[4] - m[3])}
    e = ev[2]
    n = int((e[:, 0] != 0).sum())
    clk, tag, step = e[:n, 0], e[:n, 1] >> 32, e[:n, 1] & 0xFFFFFFFF
    t10 = {int(s): int(c) for c, t, s in zip(clk, tag, step) if t == 10}
    t11 = {int(s): int(c) for c, t, s in zip(clk, tag, step) if t == 11}
    ks = sorted(t10)[16:]
    res["pv_issuer"] = {"events": n, "period_per_step": float(np.median(np.diff([t10[s] for s in ks]))),
                        "pv_issue": float(np.median([t11[s] - t10[s] for s in ks if s in t11])),
                        "wait_for_p": float(np.median([t10[s + 1] - t11[s] for s in ks if s + 1 in t10 and s in t11]))}
    e = ev[3]
    n = int((e[:, 0] != 0).sum())
    clk, step = e[:n, 0], e[:n, 1] & 0xFFFFFFFF
    t12 = {int(s): int(c) for c, s in zip(clk, step)}
    res["s_issuer"] = {"events": n, "period_per_step": float(np.median(np.diff([t12[s] for s in sorted(t12)[16:]]))),
                       "s_issued_before_p_seen": float(np.median([t10[s] - t12[s] for s in ks if s in t12]))}
    # hand-off: p_full arrive (tag 4 of softmax) -> MMA saw it (tag 10)
    print(name, json.dumps(res, indent=1))
    return res


CODES = {1: "softmax s_full", 2: "softmax pv_done (rescale)", 3: "softmax o_empty", 4: "epilogue l_ready", 5: "epilogue o_done",
         10: "tma q_empty", 11: "tma k_empty", 12: "tma v_empty", 20: "pv v_full", 21: "pv o_empty", 22: "pv p_full",
         30: "s q_full", 31: "s k_full", 32: "s pv_done"}

if __name__ == "__main__":
    which = sys.argv[1:] or ["vit", "llm"]
    out = {}
    for w in which:
        fn = run_vit() if w == "vit" else run_llm()
        # pinned HOST buffer (device-visible under UVA): records written before a watchdog trap survive the dead context
        buf = torch.zeros(4 * EV * 2 + 16 * 8, dtype=torch.int64).pin_memory()
        os.environ["MJV_ATTN_TRACE"] = hex(buf.data_ptr())
        try:
            fn()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(w, "FAILED:", repr(e)[:120])
            rec = buf[4 * EV * 2:].view(16, 8).numpy()
            for wi in range(16):
                if rec[wi, 0] != 0:
                    print(f"  warp {wi}: wedged in '{CODES.get(int(rec[wi, 0]), rec[wi, 0])}' parity {rec[wi, 1]} block {rec[wi, 2]} "
                          f"thread {rec[wi, 3]} aux item {rec[wi, 4] >> 32} low {hex(int(rec[wi, 4] & 0xffffffff))}")
            raw = buf[:4 * EV * 2].view(4, EV, 2).numpy()
            for role in range(4):
                n = int((raw[role, :, 0] != 0).sum())
                last = raw[role, max(0, n - 3):n]
                print(f"  role {role}: {n} events, last: {[(int(c), int(x >> 32), int(x & 0xffffffff)) for c, x in last]}")
            sys.exit(1)
        os.environ.pop("MJV_ATTN_TRACE")
        out[w] = summarize(buf[:4 * EV * 2].clone(), w)
        np.save(f"gpurun_out/attn_trace_{w}.npy", buf[:4 * EV * 2].view(4, EV, 2).numpy())
    with open("gpurun_out/attn_trace.json", "w") as f:
        json.dump(out, f, indent=1)
