"""GPU: the C ABI is re-entrant per stream (include/gsr_b200.h): host threads, each on its own CUDA stream, run forward +
backward of different scenes / rasterizers at the same time and must reproduce their single-threaded results -- the
forward images bit for bit (the per-thread pinned read-back slots, the per-configuration capacity history and the
caller-owned buffers are the only state), gradients to float-atomic rounding."""
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

import harness as hz  # noqa: E402
import synth  # noqa: E402


def _jobs():
    jobs = []
    for i, (P, W, H) in enumerate([(30000, 320, 240), (50000, 400, 304), (20000, 640, 360), (40000, 256, 256)]):
        sc = synth.make_scene(P, W, H, seed=40 + i)
        gc, go = synth.make_upstream_grads(W, H, seed=50 + i)
        jobs.append(("surfel", sc, gc, go))
    for i, (P, W, H) in enumerate([(30000, 320, 240), (25000, 480, 270)]):
        sc = synth.make_scene(P, W, H, seed=60 + i, scale_dims=3)
        gc, _ = synth.make_upstream_grads(W, H, seed=70 + i)
        jobs.append(("gauss", sc, gc, None))
    return jobs


def _run(job):
    kind, sc, gc, go = job
    out = hz.run_product_surfel(sc, gc, go) if kind == "surfel" else hz.run_product_gauss(sc, gc)
    torch.cuda.current_stream().synchronize()
    return out


def test_threads_on_their_own_streams_reproduce_the_serial_results():
    jobs = _jobs()
    serial = [_run(j) for j in jobs]
    rounds = 6
    results = [[None] * rounds for _ in jobs]
    errors = []

    def worker(k):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()):
                for r in range(rounds):
                    results[k][r] = _run(jobs[k])
        except Exception as e:      # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k, want in enumerate(serial):
        for r in range(rounds):
            got = results[k][r]
            assert np.array_equal(got["color"], want["color"]) and np.array_equal(got["radii"], want["radii"]), (k, r)
            if "others" in want:
                assert np.array_equal(got["others"], want["others"]), (k, r)
            for name, g in want["grads"].items():
                assert hz.rel_linf(got["grads"][name], g) <= 2e-5, (k, r, name)
