import sys, os; sys.path.insert(0,".")
from mc_mpi_b200 import configs
from mc_mpi_b200.layer import decompose_domain
def run(K,R,n,opts):
    cfg=configs.single_gpu_slab(n)
    g=decompose_domain(cfg.x_min,cfg.x_max,cfg.x_ini,K,R,cfg.nb_cells,cfg.nb_particles,cfg.particle_min_weight,global_dx=K>1)
    for k,v in opts.items(): g.set_option(k,v)
    best=0
    for rep in range(3):
        g.create_particles(cfg.x_ini,1.0/cfg.nb_particles,cfg.nb_particles); c0=g.counts(); c=g.simulate(-1); g.pop_left(); g.pop_right()
        best=max(best,(c["events"]-c0["events"])/(c["track_ms"]-c0["track_ms"])*1e3)
    g.close(); return best
lib=os.environ.get("MCB200_LIB","default")
for bps in (4,5,6):
    print(lib.split("/")[-1], "bps",bps, "full %.4e"%run(1,0,30_000_000,{"blocks_per_sm":bps}), "sub(K8,r5) %.4e"%run(8,5,50_000_000,{"blocks_per_sm":bps}), flush=True)
