"""OpenMP scaling of the C oracle port on the host (run on the GPU box)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import bench
    from vbmc_b200 import workloads
    from oracle import cport
    w = workloads.build(sys.argv[1], bench.scipy_gp_post, with_eps=False)
    print(json.dumps(cport.time_negelcbo(w, steps=2, warmup=1)))
else:
    print(open("/sys/fs/cgroup/cpu.max").read() if os.path.exists("/sys/fs/cgroup/cpu.max") else "no cpu.max")
    print("affinity", len(os.sched_getaffinity(0)))
    for t in (8, 16, 32, 64, 128):
        env = dict(os.environ, OMP_NUM_THREADS=str(t), OMP_PROC_BIND="spread")
        out = subprocess.run([sys.executable, __file__, "c3"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        print(t, out)
