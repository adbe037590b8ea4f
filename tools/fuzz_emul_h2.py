"""One-off fuzz run on the CPU (not part of the test suite): the h2 / gRPC server scenario of tests/test_gpu_h2_server.py — many connections,
byte streams delivered in random pieces over many batches, CONTINUATION / padding / trailers / interleaved streams, protocol violations —
over many seeds through the EMULATED library (tests/test_emulated_library.py) against the oracle's H2 connection state machine.
Usage: python tools/fuzz_emul_h2.py [seconds] [base seed]"""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul_runner
ctypes.CDLL = emul_runner.EmulCDLL
import brpc_b200
from _h2server_scenario import run_scenario


def run(budget, base_seed):
    t0 = time.time(); seed = 0; tm = tc = 0
    mk = lambda: brpc_b200.Context(device=0, max_batch_bytes=8 << 20, max_msgs=1 << 14, max_runs=256, max_resp_bytes=16 << 20)
    while time.time() - t0 < budget:
        seed += 1
        rng = random.Random(base_seed + seed)
        msgs, ctrl, errors = run_scenario(mk, brpc_b200.make_runs, rng.randrange(1, 24), rng.randrange(1, 16), rng.choice([0.0, 0.1, 0.25, 0.5]), base_seed + seed,
                                          rng.choice([[1, 9, 100, 1500], [3, 50, 700, 4000, 30000], [200000]]))
        tm += msgs; tc += ctrl
    return seed, tm, tc


if __name__ == "__main__":
    sd, m, c = run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 9500000)
    print("h2 fuzz ok: %d scenarios, %d gRPC messages, %d control frames, emulated library == oracle everywhere" % (sd, m, c))
