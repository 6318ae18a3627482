"""Developer helper (not a pytest): A/B whole builds of libgsr_b200.so on cfg-B.

    python tests/gpu_lib_sweep.py gs-sr_b200/variants/libgsr_r1.so gs-sr_b200/libgsr_b200.so ...

Each library runs in its own process (GSR_B200_LIB): per-kernel cudaEvent times (library profiling slots) and the
gradients' agreement with the FIRST library of the list."""
import ctypes, os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def child(ref_npz, P, W, H):
    sys.path.insert(0, HERE)
    import gpu_profile as gp
    import torch, gsr_b200
    sc, tt, gct, got, rast, leaves, m2d = gp.setup(P, W, H)
    L = gsr_b200.lib()
    names = ["pre_fwd", "scan", "dup", "sort", "build", "render_fwd", "render_bwd", "pre_bwd"]
    for _ in range(3):
        gp.product_step(rast, leaves, m2d, gct, got)
    acc = np.zeros(16); n = 6
    for _ in range(n):
        for t in leaves.values():
            t.grad = None
        L.gsr_profile_enable(1)
        gp.product_step(rast, leaves, m2d, gct, got)
        buf = (ctypes.c_float * 16)(); L.gsr_profile_read(buf); acc += np.array(list(buf))
    L.gsr_profile_enable(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gp.product_step(rast, leaves, m2d, gct, got)
    e1.record(); torch.cuda.synchronize()
    msg = " ".join(f"{k}={acc[i]/n*1e3:.0f}" for i, k in enumerate(names)) + f" | step={e0.elapsed_time(e1)/20:.3f} ms"
    for t in leaves.values():
        t.grad = None
    gp.product_step(rast, leaves, m2d, gct, got)
    g = {k: t.grad.cpu().numpy() for k, t in leaves.items()}
    if ref_npz and os.path.exists(ref_npz):
        r = np.load(ref_npz)
        msg += " | grad rel-max diff " + " ".join(f"{k}={np.abs(g[k]-r[k]).max()/np.abs(r[k]).max():.1e}" for k in g)
    elif ref_npz:
        np.savez(ref_npz, **g)
    print(msg, flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], *map(int, sys.argv[3:6]))
    else:
        libs = [a for a in sys.argv[1:] if not a.startswith("--")]
        size = [a[7:] for a in sys.argv[1:] if a.startswith("--size=")]
        P, W, H = (size[0].split(",") if size else ("2000000", "1600", "1060"))
        ref = f"/tmp/gsr_sweep_ref_{P}_{W}_{H}.npz"
        if os.path.exists(ref):
            os.remove(ref)
        for lib in libs:
            env = dict(os.environ, GSR_B200_LIB=os.path.abspath(lib))
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", ref, P, W, H], env=env,
                               capture_output=True, text=True, timeout=600)
            print(f"{os.path.basename(lib):36s} {p.stdout.strip() or p.stderr.strip()[-600:]}", flush=True)
