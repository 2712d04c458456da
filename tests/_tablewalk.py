"""TEST-ONLY model of what the CUDA kernels do with the device tables (rmqtt_b200/csrc/kernels.cuh),
written as slow pure Python over the host mirror exported by gm_debug_table.  It lets the CPU test
tier check the host-side trie builder (host_trie.cpp) and the table layout (layout.h) against the
oracle without a GPU.  It is not importable from the product package and matches nothing at run time.
"""
import numpy as np

M32 = 0xFFFFFFFF
CNT_BIG = 0xFFFF
TOK_UNKNOWN, TOK_PLUS, TOK_HASH, TOK_BLANK = 0, 1, 2, 3


def fmix32(h):
    h ^= h >> 16; h = (h * 0x85EBCA6B) & M32; h ^= h >> 13; h = (h * 0xC2B2AE35) & M32; h ^= h >> 16
    return h


def edge_hash(parent, token):
    return fmix32((parent * 0x9E3779B1 + ((token ^ 0x7F4A7C15) * 0x85EBCA77)) & M32)


def fnv(b):
    h = 0x811C9DC5
    for c in b:
        h = ((h ^ c) * 0x01000193) & M32
    return h


def pack_words(b):
    buf = bytearray(28)
    buf[:len(b)] = b
    buf[27] = len(b)
    return [int.from_bytes(buf[4 * k:4 * k + 4], "little") for k in range(7)]


def dict_hash(b):
    if len(b) <= 27:
        h = 0x811C9DC5
        for w in pack_words(b):
            h = ((h ^ w) * 0x01000193) & M32
        return fmix32(h)
    return fmix32(fnv(b) ^ ((len(b) * 0x9E3779B1) & M32))


def cfilter_pos(parent, token, word_mask):
    h = fmix32((((parent ^ 0x68E31DA4) * 0x9E3779B1) + token * 0x85EBCA77) & M32)
    g = fmix32((h + 0x9E3779B9) & M32)
    return h & word_mask, (1 << (g & 31)) | (1 << ((g >> 5) & 31))


MASK_BLOOM = 0x007FFFFF
MASK_WIDE_FLAG = 1 << 23
WTAG_SHIFT = 24


def mask_bit(tok):
    return 1 << ((((((tok * 0x9E3779B1) & M32) >> 16) * 23) & M32) >> 16)


class Tables:
    def __init__(self, t):
        self.edges, self.ranges, self.values = t["edges"], t["ranges"], t["values"]
        self.dict, self.pool = t["dict"], t["pool"]
        self.cfilter = t["cfilter"]
        (self.root_plus, self.root_hash_ref, self.root_mask, self.max_depth, self.root_hash_cnt,
         self.win_mask, self.win_shift, self.nwin_mask) = (int(x) for x in t["root"])
        assert (self.win_mask + 1) * (self.nwin_mask + 1) == len(self.edges) and self.win_mask + 1 == 1 << self.win_shift
        self.dict_bytes = self.dict.view(np.uint8).reshape(len(self.dict), 32)

    def token(self, lv: bytes):
        if lv == b"":
            return TOK_BLANK
        if lv == b"+":
            return TOK_PLUS
        if lv == b"#":
            return TOK_HASH
        mask = len(self.dict) - 1
        i = dict_hash(lv) & mask
        while True:
            tok = int(self.dict[i, 0])
            if tok == 0:
                return TOK_UNKNOWN
            row = self.dict_bytes[i]
            if len(lv) <= 27:
                if row[31] == len(lv) and bytes(row[4:4 + len(lv)]) == lv and not any(row[4 + len(lv):31]):
                    return tok
            elif row[31] == 0xFF and int(self.dict[i, 1]) == len(lv) and int(self.dict[i, 3]) == fnv(lv):
                off = int(self.dict[i, 2])
                if bytes(self.pool[off:off + len(lv)]) == lv:
                    return tok
            i = (i + 1) & mask

    def tokenize(self, topic: bytes):
        """-> (tokens, dollar) or None if invalid (mirrors k_tokenize)."""
        toks, dollar = [], False
        levels = topic.split(b"/")
        for k, lv in enumerate(levels):
            last = k == len(levels) - 1
            if lv == b"#" and not last:
                return None
            if lv not in (b"+", b"#") and (b"+" in lv or b"#" in lv):
                return None
            if lv[:1] == b"$":
                if k > 0:
                    return None
                dollar = True
            toks.append(self.token(lv))
        return toks, dollar

    def probe(self, parent, tok, parent_mask):
        """Linear probing inside the window named by the top byte of the parent's mask word (layout.h)."""
        wm = self.win_mask
        base = ((parent_mask >> WTAG_SHIFT) & self.nwin_mask) << self.win_shift
        i = base | (edge_hash(parent, tok) & wm)
        self.windows_touched.add(base)
        while True:
            e = self.edges[i]
            if e[2] == 0:
                return None
            if e[0] == parent and e[1] == tok:
                return dict(node=int(e[2]), plus=int(e[3]), hash_ref=int(e[4]), own_ref=int(e[5]), mask=int(e[6]), cnts=int(e[7]))
            i = base | ((i + 1) & wm)

    def expand(self, ref, cnt, out):
        if cnt == 0:
            return 0
        if cnt == 1:
            out.append(ref)
            return 1
        off = ref
        if cnt == CNT_BIG:
            off, cnt = (int(x) for x in self.ranges[ref])
        out.extend(int(v) for v in self.values[off:off + cnt])
        return 1

    def match(self, topic: bytes):
        """-> (sorted ids, counters dict) or (None, None)."""
        tk = self.tokenize(topic)
        if tk is None:
            return None, None
        toks, dollar = tk
        L = len(toks)
        out = []
        V = E = F = 0
        self.windows_touched = set()
        stack = [(dict(node=0, plus=self.root_plus, hash_ref=self.root_hash_ref, own_ref=0, mask=self.root_mask, cnts=self.root_hash_cnt), 0, dollar)]
        while stack:
            r, d, dollar_root = stack.pop()
            V += 1
            if not dollar_root:
                F += self.expand(r["hash_ref"], r["cnts"] & 0xFFFF, out)
            if d == L:
                F += self.expand(r["own_ref"], r["cnts"] >> 16, out)
                continue
            E += 1
            plus_idx = r["plus"]
            if plus_idx and not dollar_root:      # direct slot of the '+' child (1-based)
                p = self.edges[plus_idx - 1]
                assert int(p[0]) == r["node"] and int(p[1]) == TOK_PLUS and int(p[2]) != 0, "stale '+' slot"
                stack.append((dict(node=int(p[2]), plus=int(p[3]), hash_ref=int(p[4]), own_ref=int(p[5]), mask=int(p[6]), cnts=int(p[7])), d + 1, False))
            t = toks[d]
            maybe = True
            if r["mask"] & MASK_WIDE_FLAG:      # child filter of wide nodes: must never give a false negative
                w, bits = cfilter_pos(r["node"], t, len(self.cfilter) - 1)
                maybe = (int(self.cfilter[w]) & bits) == bits
            if t != TOK_UNKNOWN and r["mask"] & MASK_BLOOM & mask_bit(t):
                c = self.probe(r["node"], t, r["mask"])
                assert maybe or c is None, "child filter false negative"
                if not maybe:
                    c = None
                if c is not None:
                    stack.append((c, d + 1, False))
        return sorted(out), dict(V=V, E=E, F=F, M=len(out), L=L, B=len(topic))
