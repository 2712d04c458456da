"""Seeded random filter / topic generators for the differential tests (edge cases of SURVEY §8a:
Blank levels, `$` roots, literal `+`/`#` in topic names, duplicate ids, deep paths, long level strings)."""
import random

ALPHA = ["a", "b", "c", "dd", "e1", "", "x" * 30, "$SYS", "$share", "+", "#", "a+", "b#", "dev-0000001", "é", "sensor/"]


def rand_level(rng, wild_ok=True, weird=0.08):
    r = rng.random()
    if r < weird:
        return rng.choice(ALPHA)
    if wild_ok and r < weird + 0.18:
        return "+"
    return rng.choice(["a", "b", "c", "d", "e", "f", "", "x" * 30, "y" * 28, "z" * 27])


def rand_filter(rng, max_depth=6):
    n = rng.randint(1, max_depth)
    lv = [rand_level(rng) for _ in range(n)]
    r = rng.random()
    if r < 0.2:
        lv[-1] = "#"
    elif r < 0.25:
        lv.insert(rng.randrange(len(lv) + 1), "#")     # mostly invalid
    if rng.random() < 0.05:
        lv[0] = rng.choice(["$SYS", "$q"])
    return "/".join(lv)


def rand_topic(rng, max_depth=7):
    n = rng.randint(1, max_depth)
    lv = [rand_level(rng, wild_ok=rng.random() < 0.05, weird=0.05) for _ in range(n)]
    if rng.random() < 0.06:
        lv[0] = rng.choice(["$SYS", "$q"])
    return "/".join(lv)
