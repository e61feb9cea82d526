"""Generate tests/golden/ref_vectors.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Every array comes out of oracle/_ref/libmcref.so, i.e. the reference's own
src/layer.cpp + src/random.cpp compiled as they lie (oracle/Makefile), driven
single-threaded so that the float tally is the reference's sequential one.
The committed .npz is what pins oracle/mc_oracle.c on machines (the GPU box)
where the reference sources do not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from mc_mpi_b200 import configs  # noqa: E402
from oracle import pyoracle  # noqa: E402
from util import RefLayer, apply_tables, make_oracle, run_chain  # noqa: E402

CASES = {
    # name: (config, world_size, world_rank)
    "test_layer": (configs.ref_test_layer(), 1, 0),
    "default_2k": (configs.reference_default(2000), 1, 0),
    "default_K5_r3": (configs.reference_default(2000), 5, 3),
    "default_K8_r5": (configs.reference_default(2000), 8, 5),
    "thick_300": (configs.optically_thick(300), 1, 0),
    "thick_K4_r2": (configs.optically_thick(100), 4, 2),
    "absdom_2k": (configs.absorption_dominated(2000), 1, 0),
    "hetero_4096": (configs.heterogeneous(4096, 64), 1, 0),
}


def main():
    out = {}
    lib = pyoracle.ref_lib()
    import ctypes as C
    # RNG known answers (src/random.cpp), SURVEY Appendix B
    s = C.c_uint64(5127801)
    chain = np.array([lib.ref_rnd_seed(C.byref(s)) for _ in range(64)], dtype=np.uint64)
    out["rng/seed_chain_5127801"] = chain
    for name, seed in (("p0", int(chain[0])), ("one", 1), ("s30061994", 30061994)):
        s = C.c_uint64(seed)
        vals = np.array([lib.ref_rnd_real(C.byref(s)) for _ in range(64)], dtype=np.float32)
        out[f"rng/real_{name}"] = vals
        out[f"rng/real_{name}_final_seed"] = np.array([s.value], dtype=np.uint64)

    for name, (cfg, K, r) in CASES.items():
        lay = make_oracle(cfg, K, r, cls=RefLayer)
        out[f"{name}/births"] = lay.particles[:16].copy()
        out[f"{name}/dx"] = np.array([lay.dx], dtype=np.float32)
        lay.simulate(-1, 1)
        out[f"{name}/weights_absorbed"] = lay.weights_absorbed.copy()
        out[f"{name}/particles_left"] = lay.particles_left.copy()
        out[f"{name}/particles_right"] = lay.particles_right.copy()
        out[f"{name}/nb_disabled"] = np.array([lay.nb_disabled], dtype=np.int64)
        print(name, "disabled", lay.nb_disabled, "L", len(lay.particles_left), "R",
              len(lay.particles_right))

    # the sync worker's loop without MPI on the reference layers (5 ranks, config.yaml cycle)
    cfg = configs.reference_default(2000)
    layers = [make_oracle(cfg, 5, r, cls=RefLayer) for r in range(5)]

    def pop(l, side):
        arr = (l.particles_left if side == 0 else l.particles_right).copy()
        (l.clear_left if side == 0 else l.clear_right)()
        return arr

    cycles, mig = run_chain(layers, 500, cfg.nb_particles,
                            pop_left=lambda l: pop(l, 0), pop_right=lambda l: pop(l, 1),
                            push=lambda l, p: l.push(p), simulate=lambda l, n: l.simulate(n, 1),
                            disabled=lambda l: l.nb_disabled)
    out["chain_K5/cycles_migrations"] = np.array([cycles, mig], dtype=np.int64)
    out["chain_K5/weights_absorbed"] = np.concatenate([l.weights_absorbed for l in layers])
    out["chain_K5/nb_disabled"] = np.array([l.nb_disabled for l in layers], dtype=np.int64)
    print("chain_K5", cycles, mig)

    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
