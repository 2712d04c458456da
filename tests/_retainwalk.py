"""TEST-ONLY pure-Python model of k_retain_step / k_retain_expand (rmqtt_b200/csrc/retain_kernels.cuh) over
the flattened retained tree exported by gm_debug_table (7: nodes, 8: child blocks, 9: pre-order values)."""
from _tablewalk import Tables, TOK_PLUS, TOK_HASH, TOK_BLANK, TOK_UNKNOWN

RF_LIT_PLUS, RF_LIT_HASH, RF_SUB_LIT_HASH, RF_HAS_VAL = 1, 2, 4, 8


class RetainTables:
    def __init__(self, t):
        self.T = Tables(t)                      # shared dictionary
        self.nodes, self.kids, self.vals = t["rnodes"], t["rkids"], t["rvals"]
        self.edges = {}
        for n in range(len(self.nodes)):
            fk, nk = int(self.nodes[n, 0]), int(self.nodes[n, 1])
            for j in range(nk):
                self.edges[(n, int(self.kids[fk + j, 0]))] = int(self.kids[fk + j, 1])
        # root bookkeeping as flatten() computes it
        fk, nk = int(self.nodes[0, 0]), int(self.nodes[0, 1])
        self.dollar = set()

    def set_root_plain(self, plain_kids, plain_val_hi):
        self.root_plain_kids, self.root_plain_val_hi = plain_kids, plain_val_hi

    def match(self, filt: bytes):
        tk = self.T.tokenize(filt)
        if tk is None:
            return None
        toks, _ = tk
        L = len(toks)
        out = []
        frontier = [(0, 0)]
        while frontier:
            nxt = []
            for node, pos in frontier:
                fk, nk, val, vlo, vhi, flags = (int(x) for x in self.nodes[node, :6])
                if nk == 0 or pos == L:
                    if pos == L and flags & RF_HAS_VAL:
                        out.append(val)
                    continue
                tok = toks[pos]
                next_hash = pos + 1 < L and toks[pos + 1] == TOK_HASH
                exact_try = tok >= TOK_BLANK or (tok == TOK_PLUS and flags & RF_LIT_PLUS) or (tok == TOK_HASH and flags & RF_LIT_HASH)
                child = self.edges.get((node, tok), 0) if exact_try else 0
                root = node == 0
                if child:
                    if next_hash and int(self.nodes[child, 5]) & RF_HAS_VAL:
                        out.append(int(self.nodes[child, 2]))
                    nxt.append((child, pos + 1))
                elif tok == TOK_PLUS or (tok == TOK_HASH and flags & RF_SUB_LIT_HASH):
                    n = self.root_plain_kids if root else nk
                    for j in range(n):
                        ktok, kchild, kval, knk = (int(x) for x in self.kids[fk + j])
                        has_val, kn = knk >> 31, knk & 0x7FFFFFFF
                        if tok == TOK_PLUS:
                            if pos + 1 == L:
                                if has_val:
                                    out.append(kval)
                            else:
                                if next_hash and has_val:
                                    out.append(kval)
                                if kn:
                                    nxt.append((kchild, pos + 1))
                        else:
                            if has_val:
                                out.append(kval)
                            if kn:
                                nxt.append((kchild, pos))
                elif tok == TOK_HASH:
                    lo = vlo + (1 if flags & RF_HAS_VAL else 0)
                    hi = self.root_plain_val_hi if root else vhi
                    out.extend(int(x) for x in self.vals[lo:hi])
            frontier = nxt
        return sorted(out)
