"""Pipeline timeline of the persistent attention kernel (attention4.cuh): CTA 0 records (event, clock64) per warp.
  python tools/trace_attention.py [n_events]     (run on the GPU box; prints the first events of every warp relative to t0)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from visrag_b200 import _lib as L  # noqa: E402
from visrag_b200 import ops  # noqa: E402

NAMES = {1: "blk_start", 2: "S_ready", 3: "S_in_regs", 4: "Phi_prev_stored", 5: "exp_start", 6: "exp_lo_done", 7: "Pbuf_free",
         8: "Plo_stored", 9: "Phi_buf_free", 10: "Phi_stored", 11: "last_pv_done", 12: "item_done", 0x10: "QK_A", 0x11: "QK_B",
         0x20: "PV_A_lo", 0x21: "PV_A_hi", 0x22: "PV_B_lo", 0x23: "PV_B_hi"}


def main():
    show = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    S, N, nh, hd, hs = 128, 1024, 16, 72, 80
    qkv = torch.zeros(S * N, 3, nh, hs, device="cuda")
    qkv[..., :hd] = torch.randn(S * N, 3, nh, hd, device="cuda")
    qkv[:, 2, :, hd] = 1.0
    qkv = qkv.reshape(S * N, 3 * nh * hs).bfloat16()
    cu = torch.arange(0, (S + 1) * N, N, dtype=torch.int32, device="cuda")
    out = torch.zeros(S * N, nh * hd, dtype=torch.bfloat16, device="cuda")

    def run():
        ops.attention(qkv, qkv, qkv, q_col0=0, k_col0=nh * hs, v_col0=2 * nh * hs, head_stride=hs, head_dim=hd, heads=nh,
                      batch=S, cu_k=cu, max_k=N, cu_q=cu, max_q=N, causal=False, scale=hd ** -0.5, out=out, v_ones_column=True)
    for _ in range(3):
        run()
    cap = 4096
    buf = torch.zeros(12 * cap, dtype=torch.int64, device="cuda")
    L.lib().vr_attention_set_trace(buf.data_ptr(), cap)
    run()
    torch.cuda.synchronize()
    L.lib().vr_attention_set_trace(None, 0)
    b = buf.cpu().view(12, cap)
    ev = []
    for w in range(12):
        for v in b[w].tolist():
            if v == 0:
                break
            ev.append((v & 0xFFFFFFFFFFFF, w, (v >> 48) & 0xFFFF))
    ev.sort()
    t0 = ev[0][0]
    print(f"{len(ev)} events; CTA 0 total {ev[-1][0] - t0} cycles")
    for w in (0, 4, 8, 10):
        print(f"--- warp {w}" + (" (softmax tile A)" if w == 0 else " (softmax tile B)" if w == 4 else " (MMA issuer A)" if w == 8 else " (MMA issuer B)"))
        last = None
        n = 0
        for t, ww, e in ev:
            if ww != w:
                continue
            print(f"  {t - t0:8d}  (+{0 if last is None else t - last:5d})  {NAMES.get(e, hex(e))}")
            last = t
            n += 1
            if n >= show:
                break
    # steady-state statistics over all recorded blocks of warp 0: gaps between consecutive events by type
    import collections
    gaps = collections.defaultdict(list)
    for w in (0, 4):
        last_t, last_e = None, None
        for t, ww, e in ev:
            if ww != w:
                continue
            if last_e is not None:
                gaps[(NAMES.get(last_e), NAMES.get(e))].append(t - last_t)
            last_t, last_e = t, e
    print("--- mean gap between consecutive events of a softmax warp (cycles, count)")
    for k, v in sorted(gaps.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k[0]:>16s} -> {k[1]:<16s} mean {sum(v) / len(v):8.0f}  n {len(v)}  total {sum(v)}")


if __name__ == "__main__":
    main()
