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
// "par": the bulk paths that run on all host threads (host_trie.cpp insert_batch_parallel + the parallel first flush,
// retain_tree.cpp set_batch_build), forced onto small inputs; run under ASan/UBSan and under ThreadSanitizer
static int parallel_paths() {
    setenv("GM_HOST_PAR_MIN", "1", 1);
    setenv("GM_HOST_THREADS", "6", 1);
    for (int round = 0; round < 3; ++round) {
        setenv("GM_WIN_MIN_SLOTS_LOG2", round == 1 ? "3" : "12", 1);
        HostTrie t(16);
        RetainTreeHost rt(&t);
        std::vector<std::string> keep_f;
        for (int i = 0; i < 3000; ++i) { std::string f = path(true); bool ch; t.insert(f.data(), f.size(), rnd() % 50, &ch); keep_f.push_back(f); }
        t.sync();
        for (int pass = 0; pass < 2; ++pass) {       // a batch onto a trie that holds filters, then another onto the result
            std::vector<char> blob; std::vector<u32> offs{0}, vals;
            for (int i = 0; i < 30000; ++i) { std::string f = path(true); blob.insert(blob.end(), f.begin(), f.end()); offs.push_back(blob.size()); vals.push_back(rnd() % 64); }
            t.reserve(30000);
            t.insert_batch(blob.data(), offs.data(), vals.data(), vals.size());
            t.sync();
            for (int i = 0; i < 500; ++i) { const std::string& f = keep_f[rnd() % keep_f.size()]; bool ch; t.remove(f.data(), f.size(), rnd() % 50, &ch); }
            t.sync();
        }
        std::vector<char> rblob; std::vector<u32> roffs{0}, rvals;
        for (int i = 0; i < 30000; ++i) { std::string s = path(false); rblob.insert(rblob.end(), s.begin(), s.end()); roffs.push_back(rblob.size()); rvals.push_back(i); }
        const u64 set = rt.set_batch(rblob.data(), roffs.data(), rvals.data(), rvals.size());     // empty tree: the level-by-level parallel build
        rt.prepare_flush(); rt.shipped();
        for (int i = 0; i < 2000; ++i) {             // in-place edits of the image the parallel build produced
            std::string s = path(false); bool had; u32 old;
            if (rnd() % 3) rt.set(s.data(), s.size(), rnd(), &had, &old); else rt.remove(s.data(), s.size(), &had, &old);
            if (i % 499 == 0) { rt.prepare_flush(); rt.shipped(); }
        }
        rt.prepare_flush();
        printf("par round %d ok: nodes %llu values %llu retained %llu/%llu set %llu\n", round, (unsigned long long)t.nodes_size(), (unsigned long long)t.values_size(),
               (unsigned long long)rt.values_size(), (unsigned long long)rt.nodes_size(), (unsigned long long)set);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "par") return parallel_paths();
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
