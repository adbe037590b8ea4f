"""Field-by-field comparison of a device batch result against the oracle's."""
import numpy as np

MSG_FIELDS = ["run_idx", "frame_off", "body_size", "meta_size", "correlation_id", "log_id", "attachment_size",
              "compress_type", "checksum_type", "error_code", "has_bits", "protocol", "content_type", "method_idx",
              "status", "resp_len"]
RUN_FIELDS = ["consumed", "parse_error", "n_msgs", "first_msg", "preferred_proto"]


def gather(resp, off, ln):
    """Concatenate resp[off[i]:off[i]+ln[i]] for all i (vectorised)."""
    off = off.astype(np.int64); ln = ln.astype(np.int64)
    total = int(ln.sum())
    if total == 0:
        return np.zeros(0, np.uint8)
    starts = np.cumsum(ln) - ln
    idx = np.repeat(off - starts, ln) + np.arange(total, dtype=np.int64)
    return resp[idx]


def assert_same(dev, orc, what=""):
    d_rs, d_msgs, d_resp = dev[:3]
    o_rs, o_msgs, o_resp = orc[:3]
    assert len(d_rs) == len(o_rs), what
    for f in RUN_FIELDS:
        bad = np.nonzero(d_rs[f] != o_rs[f])[0]
        assert len(bad) == 0, "%s run_status.%s differs at runs %s: dev %s oracle %s" % (what, f, bad[:5], d_rs[f][bad[:5]], o_rs[f][bad[:5]])
    assert len(d_msgs) == len(o_msgs), "%s n_msgs dev %d oracle %d" % (what, len(d_msgs), len(o_msgs))
    for f in MSG_FIELDS:
        bad = np.nonzero(d_msgs[f] != o_msgs[f])[0]
        assert len(bad) == 0, "%s msgs.%s differs at %s: dev %s oracle %s (status dev %s orc %s)" % (
            what, f, bad[:5], d_msgs[f][bad[:5]], o_msgs[f][bad[:5]], d_msgs["status"][bad[:5]], o_msgs["status"][bad[:5]])
    # client-side responses (status 7) report the message in place: offsets are batch-relative and must agree
    inplace = d_msgs["status"] == 7
    assert np.array_equal(d_msgs["resp_off"][inplace], o_msgs["resp_off"][inplace]), what + " in-place response offsets differ"
    # response bytes, message by message (the device pads slots; the oracle packs tight)
    dl = np.where(inplace, 0, d_msgs["resp_len"]); ol = np.where(inplace, 0, o_msgs["resp_len"])
    dg = gather(d_resp, d_msgs["resp_off"], dl)
    og = gather(o_resp, o_msgs["resp_off"], ol)
    assert len(dg) == len(og)
    if len(dg):
        neq = np.nonzero(dg != og)[0]
        if len(neq):
            pos = int(neq[0]); ends = np.cumsum(ol.astype(np.int64))
            mi = int(np.searchsorted(ends, pos, side="right"))
            lo = int(ends[mi] - ol[mi])
            raise AssertionError("%s response bytes differ in msg %d at byte %d:\n dev %s\n orc %s" % (
                what, mi, pos - lo, bytes(dg[lo:lo + 96]).hex(), bytes(og[lo:lo + 96]).hex()))
    # layout invariants of the device response region: replies are iovec-style (one span each) and never overlap; a run's
    # replies lie inside the run's span, or — replies the fused kernel could not build in place — behind every run's span
    if len(d_msgs):
        has = (d_msgs["resp_len"] > 0) & ~inplace
        off = d_msgs["resp_off"][has].astype(np.int64); ln = d_msgs["resp_len"][has].astype(np.int64)
        order = np.argsort(off, kind="stable")
        so, sl = off[order], ln[order]
        assert np.all(so[1:] >= so[:-1] + sl[:-1]), what + " response spans overlap"
        if len(off):
            assert int((off + ln).max()) <= len(d_resp), what + " a response lies outside the resp region"
    span_end = int((d_rs["resp_off"].astype(np.int64) + d_rs["resp_bytes"]).max()) if len(d_rs) else 0
    for r in range(len(d_rs)):
        a, n = int(d_rs["first_msg"][r]), int(d_rs["n_msgs"][r])
        if n and np.any((d_msgs["resp_len"][a:a + n] > 0) & (d_msgs["status"][a:a + n] != 7)):
            m = d_msgs[a:a + n]; m = m[(m["resp_len"] > 0) & (m["status"] != 7)]
            lo, hi = int(d_rs["resp_off"][r]), int(d_rs["resp_off"][r]) + int(d_rs["resp_bytes"][r])
            inside = (m["resp_off"].astype(np.int64) >= lo) & (m["resp_off"].astype(np.int64) + m["resp_len"] <= hi)
            behind = m["resp_off"].astype(np.int64) >= span_end
            assert np.all(inside | behind), what + " run %d: a response is neither in the run's span nor in the overflow area" % r
