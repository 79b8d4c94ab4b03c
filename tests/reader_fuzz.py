"""Damaged mesh files through the product's readers (run by tests/test_reader_fuzz_cpu.py in a process of its own, with the
address space capped at 3 GB so that a count taken from a damaged header cannot size an allocation unnoticed):
    python -m tests.reader_fuzz <tau|foam> <seed> <mutations> <work dir>
Every mutation must end in a mesh or in an FjsphError that is not an allocation failure; a crash ends the process."""
import glob
import os
import pathlib
import resource
import shutil
import sys

import numpy as np

from fjsph_b200 import _lib, frontend


def mutate(b, rng, ascii_digits=False):
    b = bytearray(b)
    mode = rng.integers(0, 4 if ascii_digits else 3)
    if mode == 0:
        for _ in range(rng.integers(1, 6)):
            b[rng.integers(0, min(len(b), 2000))] = rng.integers(0, 256)     # the header region
    elif mode == 1:
        b = b[:rng.integers(4, len(b))]                                       # truncation
    elif mode == 2:
        k = rng.integers(0, len(b) - 8)
        b[k:k + 4] = rng.integers(0, 256, size=4, dtype=np.uint8).tobytes()   # four bytes anywhere
    else:
        idx = [i for i, c in enumerate(b) if 48 <= c <= 57]                   # counts and labels of an ASCII file
        for i in rng.choice(idx, size=min(3, len(idx)), replace=False):
            b[i] = 48 + rng.integers(0, 10)
    return bytes(b)


def main():
    kind, seed, count, root = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), pathlib.Path(sys.argv[4])
    resource.setrlimit(resource.RLIMIT_AS, (3 << 30, 3 << 30))
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-0.1, -0.1, -0.1]), np.array([0.1, 0.1, 0.1])
    read = errors = 0

    def attempt(f):
        nonlocal read, errors
        try:
            f()
            read += 1
        except _lib.FjsphError as e:
            errors += 1
            assert "alloc" not in str(e) and "length_error" not in str(e), str(e)

    if kind == "tau":
        from tests.tau_case import write_tau, write_tau_edge

        (root / "e").mkdir(parents=True, exist_ok=True)
        mesh, sol, *_ = write_tau(root, lo, hi, (3, 3, 2), lambda x: (1.0, 2.0, 3.0), lambda x: 1e5, lambda x: 1.2)
        emesh, esol, *_ = write_tau_edge(root / "e", lo[:2], hi[:2], (4, 3), lambda x: (1.0, 2.0), lambda x: 1e5, lambda x: 1.2)
        for it in range(count):
            which = it % 4
            dst = str(root / ("mut%d" % which))
            open(dst, "wb").write(mutate(open([mesh, sol, emesh, esol][which], "rb").read(), rng))
            attempt([lambda: frontend.read_tau(dst, sol), lambda: frontend.read_tau(mesh, dst),
                     lambda: frontend.read_tau_edge(dst, esol, offset_axis=2),
                     lambda: frontend.read_tau_edge(emesh, dst, offset_axis=2)][which])
    else:
        from tests.foam_case import write_case

        base = {}
        for binary in (False, True):
            base[binary] = root / ("foam_%d" % binary)
            write_case(base[binary], lo, hi, (3, 3, 2), lambda c: (1.0, 2.0, 3.0), lambda c: 1e5, binary=binary)
        files = {b: sorted(f for f in glob.glob(str(base[b] / "**" / "*"), recursive=True) if os.path.isfile(f)) for b in base}
        for it in range(count):
            binary = bool(it % 2)
            work = root / "work"
            if work.exists():
                shutil.rmtree(work)
            shutil.copytree(base[binary], work)
            victim = files[binary][rng.integers(0, len(files[binary]))]
            open(work / os.path.relpath(victim, base[binary]), "wb").write(mutate(open(victim, "rb").read(), rng, ascii_digits=True))
            attempt(lambda: frontend.read_foam(str(work), "100"))
    print("%s: %d mutations, %d still read, %d reported as errors" % (kind, count, read, errors))


if __name__ == "__main__":
    main()
