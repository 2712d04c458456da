// ASan/UBSan stress of the host-side builders (no oracle here: memory safety and internal consistency only).
#include "retain_tree.h"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
using namespace gm;
static unsigned long long x = 88172645463325252ull;
static unsigned long long rnd() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; }
static std::string path(bool filter) {
    static const char* lv[] = {"a", "b", "c", "dd", "", "$SYS", "+", "#", "a+", "xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx", "dev-0000001", "e", "f"};
    int n = 1 + rnd() % 6; std::string s;
    for (int i = 0; i < n; ++i) { if (i) s += '/'; int k = rnd() % 13; if (!filter && (k == 6 || k == 7) && rnd() % 4) k = 0; s += lv[k]; }
    return s;
}
int main() {
    for (int round = 0; round < 4; ++round) {
        setenv("GM_WIN_MIN_SLOTS_LOG2", round % 2 ? "3" : "12", 1);
        HostTrie t(16);
        RetainTreeHost rt(&t);
        std::vector<std::string> fs, ts;
        std::vector<char> blob; std::vector<u32> offs{0}, vals;
        for (int i = 0; i < 20000; ++i) {
            int op = rnd() % 10;
            if (op < 4) { std::string f = path(true); bool ch; t.insert(f.data(), f.size(), rnd() % 50, &ch); fs.push_back(f); }
            else if (op < 6 && !fs.empty()) { const std::string& f = fs[rnd() % fs.size()]; bool ch; t.remove(f.data(), f.size(), rnd() % 50, &ch); }
            else if (op < 8) { std::string s = path(false); bool had; u32 old; rt.set(s.data(), s.size(), rnd(), &had, &old); ts.push_back(s); }
            else if (!ts.empty()) { const std::string& s = ts[rnd() % ts.size()]; bool had; u32 old; rt.remove(s.data(), s.size(), &had, &old); }
            if (i % 997 == 0) { t.sync(); rt.prepare_flush(); rt.shipped(); }
            if (i % 7001 == 7000) { std::vector<u32> keep = rt.used_tokens(), remap; t.compact(&keep, &remap); rt.remap_tokens(remap); }
        }
        for (int i = 0; i < 5000; ++i) { std::string f = path(true); blob.insert(blob.end(), f.begin(), f.end()); offs.push_back(blob.size()); vals.push_back(i); }
        t.reserve(5000);
        u64 ch = t.insert_batch(blob.data(), offs.data(), vals.data(), vals.size());
        t.sync(); rt.prepare_flush();
        printf("round %d ok: nodes %llu values %llu retained %llu/%llu changed %llu\n", round, (unsigned long long)t.nodes_size(), (unsigned long long)t.values_size(),
               (unsigned long long)rt.values_size(), (unsigned long long)rt.nodes_size(), (unsigned long long)ch);
    }
    return 0;
}
