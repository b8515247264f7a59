"""bench.py's reference arm (CPU): uses every host thread whatever OMP_NUM_THREADS says (torchrun sets it to 1), prints
the same `config` object as the CUDA arm, and never maps libgrr_b200.so (the problem is described with a lazy chain)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_threads_config_and_no_product_library():
    code = (
        "import sys, json, io, contextlib\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-seconds', '1.5', '--gpus', '2']\n"
        "import bench\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf):\n"
        "    bench.main()\n"
        "from globalredundancyresolution_b200 import _cabi\n"
        "maps = open('/proc/self/maps').read()\n"
        "line = json.loads(buf.getvalue().strip().splitlines()[-1])\n"
        "line['_lib_loaded'] = _cabi._lib is not None or 'libgrr_b200' in maps\n"
        "print(json.dumps(line))\n")
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["_lib_loaded"] is False
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    sys.path.insert(0, ROOT)
    import bench
    import globalredundancyresolution_b200 as grr
    pdef = grr.problems.read_options("planar_20R", "baseline_cfg2.json", {"Nw": 20000})
    # weak scaling at 2 GPUs: 2 x the nodes; the CUDA arm builds its config with the same function and arguments
    pdef2, res = grr.load_problem("planar_20R", "baseline_cfg2.json", overrides={"Nw": 20000}, lazy_chain=True)
    res.sampleWorkspace(20000, method="staggered_grid")
    assert line["config"] == bench.config_dict(pdef, res.ws_params.shape[0], res.ws_edges.shape[0], 2)


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
