"""GPU diagnostic: role timeline of CTA 0 of one implicit-GEMM launch (diagnostic build with -DB2U_TIMELINE).

usage: python tools/conv_timeline.py [CASE=res100]      (cases of tools/one_conv.py)
Builds unet_b200/_obj_tl/libb2u_tl.so from the same sources with the timeline stamps compiled in, runs the case twice and
prints, per tile of CTA 0: when the producer got a free A stage, when the MMA issuer received its A stages, when it
committed the accumulator, when the epilogue received it and when it finished storing - all in SM clocks relative to
the first stamp - plus the mean per-tile period of every role.  The product library never contains the stamps."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OBJ = os.path.join(ROOT, "unet_b200", "_obj_tl")
LIB = os.path.join(OBJ, "libb2u_tl.so")


def build():
    from unet_b200.build import NVCC_FLAGS, sources
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for src in sources():
        o = os.path.join(OBJ, src.stem + ".o")
        subprocess.run(["nvcc", *NVCC_FLAGS, "-DB2U_TIMELINE", "-c", str(src), "-o", o], check=True)
        objs.append(o)
    subprocess.run(["nvcc", "-shared", "-o", LIB, *objs], check=True)


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "res100"
    if not os.path.exists(LIB) or "--rebuild" in sys.argv:
        build()
    os.environ["B2U_LIB"] = LIB
    import ctypes as C
    import torch
    from unet_b200 import _lib, ops
    sys.argv = ["one_conv.py", case, "1"]
    ns = {}
    src = open(os.path.join(ROOT, "tools", "one_conv.py")).read()
    ns["__file__"] = os.path.join(ROOT, "tools", "one_conv.py")
    exec(compile(src, "one_conv.py", "exec"), ns)          # builds the plan and runs it (warm-up + timed)
    plan = ns["plan"]
    lib = _lib.load()
    EV = 4096
    buf = torch.zeros(4 * EV, dtype=torch.int64, device="cuda")
    lib.b2u_conv_plan_set_timeline.argtypes = [C.c_void_p, C.c_void_p]
    _lib.check(lib.b2u_conv_plan_set_timeline(plan.handle, C.c_void_p(buf.data_ptr())), "set_timeline")
    plan.run()
    torch.cuda.synchronize()
    t = buf.cpu().view(4, EV)
    names = ["producer: A stage free", "mma: A stage arrived", "mma: accumulator committed", "epilogue: acc arrived / chunk stored"]
    t0 = int(t[t > 0].min())
    for r in range(4):
        v = t[r][t[r] > 0] - t0
        if len(v) > 8:
            d = (v[1:] - v[:-1]).float()
            print(f"{names[r]:38s} events {len(v):5d}  mean period {d[4:].mean():8.1f} clk  median {d[4:].median():8.1f}  max {d[4:].max():8.0f}")
    kc = plan.info.k_chunks
    per_tile_mma = t[2][t[2] > 0] - t0
    print("first 12 tiles (clk since start): accumulator committed", [int(x) for x in per_tile_mma[:12]])
    ep = t[3][t[3] > 0] - t0
    print("first 24 epilogue stamps:", [int(x) for x in ep[:24]])
    print("epilogue stamps 200..260 (deltas):", [int(ep[i + 1] - ep[i]) for i in range(200, min(260, len(ep) - 1))])
    print("plan: SA", plan.info.stages, "info", [(n, getattr(plan.info, n)) for n, _ in plan.info._fields_])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    torch.save(t, os.path.join(ROOT, "gpurun_out", f"conv_timeline_{case}.pt"))
