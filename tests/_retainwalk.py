"""TEST-ONLY pure-Python model of k_retain_step / k_retain_expand (rmqtt_b200/csrc/retain_kernels.cuh) over
the flattened retained tree exported by gm_debug_table (7: nodes, 8: child blocks, 9: pre-order values)."""
from _tablewalk import Tables, TOK_PLUS, TOK_HASH, TOK_BLANK, TOK_UNKNOWN

RF_LIT_PLUS, RF_LIT_HASH, RF_SUB_LIT_HASH, RF_HAS_VAL = 1, 2, 4, 8


class RetainTables:
    """Items carry the node's record {node, first_kid, nk_flags, val, val_lo, val_hi} exactly like the kernel's RItem;
    records come from the child-block entries (8 words: token, child, first_kid, nk_flags, val, val_lo, val_hi, pad)."""

    def __init__(self, t):
        self.T = Tables(t)                      # shared dictionary
        self.nodes, self.kids, self.vals = t["rnodes"], t["rkids"], t["rvals"]
        self.edges = {}
        for n in range(len(self.nodes)):
            fk, nk = int(self.nodes[n, 0]), int(self.nodes[n, 1])
            for j in range(nk):
                k = self.kids[fk + j]
                self.edges[(n, int(k[0]))] = tuple(int(x) for x in k[1:7])      # (child, first_kid, nk_flags, val, val_lo, val_hi)
                c = self.nodes[int(k[1])]
                # the copied record must equal the child's own record
                assert (int(k[2]), int(k[3]) & 0x0FFFFFFF, int(k[3]) >> 28, int(k[5]), int(k[6])) == (int(c[0]), int(c[1]), int(c[5]), int(c[3]), int(c[4]))

        self.check_hash(t.get("redges"))

    def check_hash(self, redges):
        """The (parent, token) hash the exact steps of the kernel probe must hold exactly the child-block entries,
        each with the same record (the kernels read records from either copy)."""
        if redges is None or len(redges) == 0:
            return
        mask = len(redges) - 1
        live = {(int(e[0]), int(e[1])): tuple(int(x) for x in e[2:8]) for e in redges if int(e[2]) != 0}
        for key, rec in self.edges.items():
            assert live.get(key) == rec, ("hash slot disagrees with the child block", key, live.get(key), rec)
            n, tok = key
            i = self.redge_hash(n, tok) & mask               # reachable by linear probing from its home slot
            while not (int(redges[i][0]) == n and int(redges[i][1]) == tok and int(redges[i][2]) != 0):
                assert int(redges[i][2]) != 0, ("probe hits an empty slot before the entry", key)
                i = (i + 1) & mask
        # entries of relocated (garbage) blocks never linger in the hash: one slot per live child entry
        assert len(live) == len(self.edges), (len(live), len(self.edges))

    @staticmethod
    def redge_hash(parent, token):
        from _tablewalk import fmix32, M32
        return fmix32(((parent * 0x85EBCA77) + ((token ^ 0x2545F491) * 0x9E3779B1)) & M32)

    def set_root_plain(self, plain_kids, plain_val_hi):
        self.root_plain_kids, self.root_plain_val_hi = plain_kids, plain_val_hi

    def match(self, filt: bytes):
        tk = self.T.tokenize(filt)
        if tk is None:
            return None
        toks, _ = tk
        L = len(toks)
        out = []
        root = (0, int(self.nodes[0, 0]), int(self.nodes[0, 1]) | (int(self.nodes[0, 5]) << 28), 0, 0, self.root_plain_val_hi)
        frontier = [(root, 0)]
        while frontier:
            nxt = []
            for (node, fk, nkf, val, vlo, vhi), pos in frontier:
                nk, flags = nkf & 0x0FFFFFFF, nkf >> 28
                if nk == 0 or pos == L:
                    if pos == L and flags & RF_HAS_VAL:
                        out.append(val)
                    continue
                tok = toks[pos]
                next_hash = pos + 1 < L and toks[pos + 1] == TOK_HASH
                exact_try = tok >= TOK_BLANK or (tok == TOK_PLUS and flags & RF_LIT_PLUS) or (tok == TOK_HASH and flags & RF_LIT_HASH)
                child = self.edges.get((node, tok)) if exact_try else None
                is_root = node == 0
                if child:
                    if next_hash and (child[2] >> 28) & RF_HAS_VAL:
                        out.append(child[3])
                    nxt.append((child, pos + 1))
                elif tok == TOK_PLUS or (tok == TOK_HASH and flags & RF_SUB_LIT_HASH):
                    n = self.root_plain_kids if is_root else nk
                    for j in range(n):
                        k = tuple(int(x) for x in self.kids[fk + j][1:7])
                        has_val, kn = (k[2] >> 28) & RF_HAS_VAL, k[2] & 0x0FFFFFFF
                        if tok == TOK_PLUS:
                            if pos + 1 == L:
                                if has_val:
                                    out.append(k[3])
                            else:
                                if next_hash and has_val:
                                    out.append(k[3])
                                if kn:
                                    nxt.append((k, pos + 1))
                        else:
                            if has_val:
                                out.append(k[3])
                            if kn:
                                nxt.append((k, pos))
                elif tok == TOK_HASH:
                    lo = vlo + (1 if flags & RF_HAS_VAL else 0)
                    hi = self.root_plain_val_hi if is_root else vhi
                    out.extend(int(x) for x in self.vals[lo:hi])
            frontier = nxt
        return sorted(out)
