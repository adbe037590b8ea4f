"""The multi-batch server-side h2 scenario shared by tests/test_gpu_h2_server.py (the device, through the C ABI) and
tests/test_device_h2_host.py (the same device source built for the host): many connections, byte streams delivered in random pieces,
every run status / control byte / message compared with the oracle."""
import random

import numpy as np

import _oracle as O
import _h2traffic as T


def _msg_tuple(m, blob, data=None):
    """data: the batch input — a device message whose body was one DATA frame points into it (B2_H2_FLAG_BODY_IN_INPUT)."""
    src = data if (int(m["flags"]) & 16) else blob
    return (int(m["stream_id"]), int(m["n_headers"]), bytes(blob[m["headers_off"]:m["headers_off"] + m["headers_len"]]),
            bytes(src[m["body_off"]:m["body_off"] + m["body_len"]]), int(m["http_method"]), int(m["content_type"]), int(m["flags"]) & ~16,
            int(m["method_idx"]), bytes(src[m["msg_off"]:m["msg_off"] + m["msg_len"]]), bytes(blob[m["path_off"]:m["path_off"] + m["path_len"]]))


def run_scenario(make_ctx, make_runs, n_conns, n_calls, violations, seed, step_choices):
    rng = random.Random(seed)
    ctx = make_ctx()
    streams = [b"".join(T.connection_script(rng, n_calls=n_calls, violations=violations if i % 2 else 0.0)) for i in range(n_conns)]
    orc = [O.H2Conn() for _ in range(n_conns)]
    for i in range(n_conns):
        ctx.h2_conn_reset(i)
    fed = [0] * n_conns; buf = [b""] * n_conns; alive = [True] * n_conns
    total_msgs = 0; total_ctrl = 0; errors = {}
    while any(alive[i] and (fed[i] < len(streams[i])) for i in range(n_conns)):
        live = [i for i in range(n_conns) if alive[i] and fed[i] < len(streams[i])]
        batch = [i for i in live if rng.random() < 0.8] or live[:1]
        for i in batch:
            k = rng.choice(step_choices)
            buf[i] += streams[i][fed[i]:fed[i] + k]; fed[i] += k
        data, runs = make_runs([buf[i] for i in batch])
        runs["socket_id"] = np.array(batch, dtype=np.uint64)
        rs, msgs, out = ctx.h2_process_batch(data, runs, out_cap=len(batch) * (512 << 10))
        for j, i in enumerate(batch):
            e, cons, omsgs, octrl, oblob, mfs, sws = orc[i].consume(buf[i])
            st = rs[j]
            assert (int(st["parse_error"]), int(st["consumed"]), int(st["n_msgs"])) == (e, cons, len(omsgs)), (i, fed[i])
            assert bytes(out[st["ctrl_off"]:st["ctrl_off"] + st["ctrl_len"]]) == octrl, (i, fed[i])
            assert (int(st["remote_max_frame_size"]), int(st["remote_stream_window_size"])) == (mfs, sws)
            dm = msgs[st["first_msg"]:st["first_msg"] + st["n_msgs"]]
            for a, b in zip(dm, omsgs):
                assert int(a["run_idx"]) == j
                assert _msg_tuple(a, out, data) == _msg_tuple(b, oblob), (i, fed[i], int(b["stream_id"]))
            total_msgs += len(omsgs); total_ctrl += len(octrl)
            buf[i] = buf[i][cons:]
            if e != 2:
                alive[i] = False; errors[e] = errors.get(e, 0) + 1
    return total_msgs, total_ctrl, errors


