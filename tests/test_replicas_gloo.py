"""The N>1 host logic on CPU: world_size 2 over gloo (127.0.0.1 rendezvous).  Replicas
partition Philox streams, the job throughput is sum(units) / max(time)."""

import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import json, sys, time
    sys.path.insert(0, %r)
    from client_b200.perf.replicas import Replicas
    r = Replicas(backend="gloo")
    r.barrier()
    units = 1000 * (r.rank + 1)            # rank 0: 1000 requests, rank 1: 2000
    seconds = 0.5 if r.rank == 0 else 2.0  # rank 1 is the slow one
    out = {"rank": r.rank, "world": r.world, "thr": r.throughput(units, seconds), "max": r.max(seconds),
           "sum": r.sum(units), "stream_base": r.stream_base(64), "seed": r.seed(7)}
    r.barrier()
    r.close()
    print("RESULT " + json.dumps(out), flush=True)
    """
)


def test_two_replicas_over_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        text, _ = p.communicate(timeout=180)
        assert p.returncode == 0, text
        line = [l for l in text.splitlines() if l.startswith("RESULT ")][0]
        import json

        outs.append(json.loads(line[7:]))
    outs.sort(key=lambda o: o["rank"])
    for o in outs:
        assert o["world"] == 2 and o["sum"] == 3000.0 and o["max"] == 2.0 and o["thr"] == 1500.0
    assert outs[0]["stream_base"] != outs[1]["stream_base"] and outs[1]["stream_base"] - outs[0]["stream_base"] >= 1 << 40
    assert outs[0]["seed"] != outs[1]["seed"]


def test_single_process_is_identity():
    from client_b200.perf.replicas import Replicas

    env = {k: os.environ.pop(k, None) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK")}
    try:
        r = Replicas()
        assert r.world == 1 and r.max(3.5) == 3.5 and r.sum(2) == 2 and r.throughput(10, 2.0) == 5.0
        r.barrier()
        r.close()
    finally:
        for k, v in env.items():
            if v is not None:
                os.environ[k] = v


WALK_WORKER = textwrap.dedent(
    """
    import json, sys, time
    sys.path.insert(0, %r)
    from client_b200.perf.replicas import Replicas
    r = Replicas(backend="gloo")
    log = []

    def good(tag, delay):
        def fn(sync):
            time.sleep(delay * (r.rank + 1))   # ranks get ready at different times
            sync()                             # ... and are released together
            log.append((tag, time.time()))
            return tag
        return fn

    def bad_on_rank1(sync):
        if r.rank == 1:
            raise RuntimeError("server did not start")   # before the rendezvous: the walk must stand in for it
        sync()
        return "b"

    results, error = r.walk([("a", good("a", 0.05)), ("b", bad_on_rank1), ("c", good("c", 0.0))])
    skipped, err2 = r.walk([("d", good("d", 0.0))], enabled=(r.rank == 0))   # one rank has nothing to run
    r.barrier()
    r.close()
    print("RESULT " + json.dumps({"rank": r.rank, "results": results, "error": error, "skipped": skipped, "err2": err2,
                                  "t_a": dict(log)["a"]}), flush=True)
    """
)


def test_plan_walk_keeps_ranks_in_step_when_one_fails(tmp_path):
    """bench.py's loopback plan (Replicas.walk): a rank that fails before the rendezvous of an entry, or has
    nothing to run, still meets the others at every barrier -- nobody hangs, later entries are skipped only there."""
    import json

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "walk_worker.py"
    script.write_text(WALK_WORKER % ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        text, _ = p.communicate(timeout=180)
        assert p.returncode == 0, text
        outs.append(json.loads([l for l in text.splitlines() if l.startswith("RESULT ")][0][7:]))
    outs.sort(key=lambda o: o["rank"])
    assert outs[0]["results"] == {"a": "a", "b": "b", "c": "c"} and outs[0]["error"] is None
    assert outs[1]["results"] == {"a": "a"} and outs[1]["error"].startswith("b: RuntimeError")
    assert outs[0]["skipped"] == {"d": "d"} and outs[1]["skipped"] == {} and outs[1]["err2"] is None
    assert abs(outs[0]["t_a"] - outs[1]["t_a"]) < 0.04  # released together although rank 1 was ready 50 ms later
