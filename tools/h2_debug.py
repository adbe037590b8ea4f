import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import brpc_b200, _oracle as O, _h2traffic as T
rng = random.Random(20260921)
msg_len=4096; K=8; n_conns=256
ctx = brpc_b200.Context(device=0, max_batch_bytes=64 << 20, max_msgs=1 << 16, max_runs=1024, max_resp_bytes=128 << 20)
message = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz0123456789") for _ in range(msg_len))
fbs=[]; bs=[]
for cidx in range(n_conns):
    ctx.h2_conn_reset(cidx)
    enc = T.HpackEncoder(rng)
    warm = b"".join(T.request_frames(rng, enc, 1, message=message, chunk=16384))
    fbs.append(T.PREFACE + T.settings() + warm)
    calls = [T.request_frames(rng, enc, 3 + 2 * k, message=message, chunk=16384) for k in range(K)]
    bs.append(b"".join(b"".join(c) for c in calls))
data0, runs0 = brpc_b200.make_runs(fbs)
rs, msgs, out = ctx.h2_process_batch(data0, runs0)
print("first", int(rs["n_msgs"].sum()), np.unique(rs["parse_error"]))
data, runs = brpc_b200.make_runs(bs)
rs, msgs, out = ctx.h2_process_batch(data, runs, msg_cap=n_conns * (K + 2), out_cap=n_conns * (K * (msg_len + 1024) * 2 + 8192))
print("second", len(msgs), np.unique(rs["parse_error"], return_counts=True), np.unique(rs["n_msgs"], return_counts=True))
badr=[i for i in range(n_conns) if rs["n_msgs"][i]!=K][:3]
for i in badr:
    c=O.H2Conn(); c.consume(fbs[i]); r=c.consume(bs[i])
    print(i, rs[i], "oracle:", r[0], r[1], len(r[2]), "ctrl dev", bytes(out[rs[i]["ctrl_off"]:rs[i]["ctrl_off"]+rs[i]["ctrl_len"]]).hex(), "ctrl orc", r[3].hex())
