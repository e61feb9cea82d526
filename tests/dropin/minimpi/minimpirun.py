#!/usr/bin/env python
"""tests/dropin/minimpi/minimpirun.py -- TEST INFRASTRUCTURE: `mpirun -n K prog args...` for minimpi.

    python tests/dropin/minimpi/minimpirun.py -n 2 [--timeout S] [--cwd DIR] prog config.yaml sync

Starts K copies of `prog` sharing one /dev/shm file (MINIMPI_SHM / MINIMPI_RANK / MINIMPI_SIZE;
LOCAL_RANK = rank so a GPU build takes one device per rank), waits for all of them, and on the
first failure or on the timeout terminates the ranks it started (by their own pids).
Exit status: 0 if every rank exited 0, else the first non-zero status (124 on timeout).
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time


def launch(n, argv, timeout=600.0, cwd=None, env=None, capture=False):
    """Run `argv` as n ranks; returns (status, stdout_of_rank0 or None)."""
    base = dict(os.environ if env is None else env)
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    fd, shm = tempfile.mkstemp(prefix="minimpi.", dir=shm_dir)
    os.close(fd)
    procs = []
    try:
        for r in range(n):
            e = dict(base, MINIMPI_SHM=shm, MINIMPI_RANK=str(r), MINIMPI_SIZE=str(n))
            e.setdefault("LOCAL_RANK", str(r))
            e.setdefault("MINIMPI_TIMEOUT_S", str(int(timeout)))
            out = subprocess.PIPE if (capture and r == 0) else None
            procs.append(subprocess.Popen(argv, cwd=cwd, env=e, stdout=out))
        deadline = time.time() + timeout
        status = 0
        live = set(range(n))
        while live and status == 0:
            for r in sorted(live):
                rc = procs[r].poll()
                if rc is not None:
                    live.discard(r)
                    if rc != 0 and status == 0:
                        status = rc
            if time.time() > deadline:
                status = 124
            if live and status == 0:
                time.sleep(0.01)
        for r in live:                      # a failed run: stop what we started
            procs[r].terminate()
        for r in live:
            try:
                procs[r].wait(timeout=5)
            except subprocess.TimeoutExpired:
                procs[r].kill()
        text = None
        if capture:
            text = procs[0].stdout.read().decode() if procs[0].stdout else ""
        return status, text
    finally:
        for p in procs:
            if p.stdout:
                p.stdout.close()
        try:
            os.unlink(shm)
        except OSError:
            pass


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("-n", type=int, default=1)
    ap.add_argument("--timeout", type=float, default=600.0)
    ap.add_argument("--cwd", default=None)
    ap.add_argument("prog", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    if not a.prog:
        ap.error("no program given")
    status, _ = launch(a.n, a.prog, timeout=a.timeout, cwd=a.cwd)
    sys.exit(status if 0 <= status < 256 else 1)


if __name__ == "__main__":
    main()
