"""CPU tier: the deterministic iot6 generator."""
import numpy as np

from rmqtt_b200 import workload as wl


def test_deterministic_and_order_free():
    a = wl.unpack(*wl.gen_subs(wl.C3, 1000)[:2])
    b = wl.unpack(*wl.gen_subs(wl.C3, 500, first=500)[:2])
    assert a[500:] == b
    assert a == wl.unpack(*wl.gen_subs(wl.C3, 1000)[:2])


def test_mix_fractions():
    subs = wl.unpack(*wl.gen_subs(wl.C3, 200_000)[:2])
    plus = sum(b"+" in s for s in subs) / len(subs)
    hsh = sum(s.endswith(b"#") for s in subs) / len(subs)
    root_plus = sum(s.startswith(b"+/") for s in subs) / len(subs)
    assert abs(plus - 0.30) < 0.01 and abs(hsh - 0.05) < 0.005
    assert abs(root_plus - 0.30 * 0.02) < 0.002
    assert all(s.count(b"/") == 5 for s in subs if not s.endswith(b"#"))


def test_c1_has_no_wildcards_and_half_hits():
    subs = set(wl.unpack(*wl.gen_subs(wl.C1)[:2]))
    assert not any(b"+" in s or b"#" in s for s in subs)
    topics = wl.unpack(*wl.gen_topics(wl.C1))
    assert len(topics) == 10_000
    hit = sum(t in subs for t in topics) / len(topics)
    assert 0.5 < hit < 0.9


def test_retained_topics_are_distinct():
    t = wl.unpack(*wl.gen_retained(wl.C4, 100_000)[:2])
    assert len(set(t)) == len(t)


def test_region_restriction():
    b, o = wl.gen_topics(wl.C3, 2000, regions=[3, 9])
    names = {t.split(b"/")[0] for t in wl.unpack(b, o)}
    assert names == {wl.region_name(3), wl.region_name(9)}


def test_zipf_over_devices_is_skewed_and_deterministic():
    b, o = wl.gen_topics_zipf(wl.C3, 50_000)
    t = wl.unpack(b, o)
    assert t == wl.unpack(*wl.gen_topics_zipf(wl.C3, 50_000))
    from collections import Counter
    c = Counter(x.split(b"/")[2] for x in t)
    top = c.most_common(1)[0][1]
    assert top > 0.02 * len(t) and len(c) > 5_000          # one hot device, a long tail
    assert all(x.count(b"/") == 5 for x in t)


def test_big_batches_are_stitched_from_per_thread_shares_with_the_same_bytes():
    """Above 200 K items the generators run on all host threads (workload.cpp `stitched`): every thread generates a contiguous
    share, the shares are concatenated at their byte offsets.  Items are order-free, so the result must be the bytes of the
    same range generated in two halves (each below the threshold: the serial loop)."""
    cfg = wl.C3.scaled(n_subs=260_000, n_topics=260_000)
    blob, offs, vals = wl.gen_subs(cfg)
    b0, o0, v0 = wl.gen_subs(cfg, n=130_000)
    b1, o1, v1 = wl.gen_subs(cfg, n=130_000, first=130_000)
    assert np.array_equal(blob, np.concatenate([b0, b1]))
    assert np.array_equal(offs, np.concatenate([o0[:-1], o1 + o0[-1]]))
    assert np.array_equal(vals, np.concatenate([v0, v1]))
    tb, to = wl.gen_topics(cfg, stream=2)
    t0, u0 = wl.gen_topics(cfg, n=130_000, stream=2)
    t1, u1 = wl.gen_topics(cfg, n=130_000, first=130_000, stream=2)
    assert np.array_equal(tb, np.concatenate([t0, t1])) and np.array_equal(to, np.concatenate([u0[:-1], u1 + u0[-1]]))
    sb, so, sv = wl.gen_subs_sharded(cfg, [1, 5, 9, 33])
    s0, p0, w0 = wl.gen_subs_sharded(cfg, [1, 5, 9, 33], n=130_000)
    s1, p1, w1 = wl.gen_subs_sharded(cfg, [1, 5, 9, 33], n=130_000, first=130_000)
    assert np.array_equal(sb, np.concatenate([s0, s1])) and np.array_equal(sv, np.concatenate([w0, w1]))
    assert np.array_equal(so, np.concatenate([p0[:-1], p1 + p0[-1]]))
