// Offline prototype (CPU, no product code): how many tree nodes does the broadphase packet walk of k_lbvh_pairs visit per packet
//   (a) on the Karras radix tree it uses today, (b) on an implicit balanced tree over the same Morton-sorted leaves (no arrival-counter
//   refit needed on the device), with 32 or 64 leaves per packet?  Same pruning rule as the kernel: a subtree is entered when some
//   query of the packet overlaps its box and the subtree holds a leaf sorted after that query (all colliders here are dynamic; the
//   terrain sits on the big-static side list).  Input: tools/proto/make_aabbs.py.   g++ -O2 -std=c++17 tree_visits.cpp -o tree_visits
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

struct Box { float mn[3], mx[3]; };
static inline bool ov(const Box& a, const Box& b) {
    for (int k = 0; k < 3; ++k) if (a.mx[k] < b.mn[k] || a.mn[k] > b.mx[k]) return false;
    return true;
}
static inline Box uni(const Box& a, const Box& b) {
    Box r;
    for (int k = 0; k < 3; ++k) { r.mn[k] = std::min(a.mn[k], b.mn[k]); r.mx[k] = std::max(a.mx[k], b.mx[k]); }
    return r;
}
static inline uint32_t expandBits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
// child >= 0: internal node, < 0: leaf ~(sorted position)
struct Tree { std::vector<int> left, right, lastL, lastR; std::vector<Box> boxL, boxR; int root = 0; };

static int n;
static std::vector<Box> leaf;        // sorted order
static std::vector<uint32_t> key;

static inline int delta(int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = key[i], b = key[j];
    if (a == b) return 32 + __builtin_clz((uint32_t)i ^ (uint32_t)j);
    return __builtin_clz(a ^ b);
}

static Box refit(Tree& t, int node) {     // iterative post-order
    struct F { int node; int stage; };
    std::vector<F> st; st.push_back({ node, 0 });
    std::vector<Box> ret;                 // value stack
    while (!st.empty()) {
        F& f = st.back();
        int c = f.stage == 0 ? t.left[f.node] : t.right[f.node];
        if (f.stage < 2) {
            ++f.stage;
            if (c < 0) ret.push_back(leaf[~c]); else st.push_back({ c, 0 });
        } else {
            Box r = ret.back(); ret.pop_back();
            Box l = ret.back(); ret.pop_back();
            t.boxL[f.node] = l; t.boxR[f.node] = r;
            ret.push_back(uni(l, r));
            st.pop_back();
        }
    }
    return ret.back();
}

static Tree karras() {
    Tree t; int m = n - 1;
    t.left.resize(m); t.right.resize(m); t.lastL.resize(m); t.lastR.resize(m); t.boxL.resize(m); t.boxR.resize(m);
    for (int i = 0; i < m; ++i) {
        int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
        int dmin = delta(i, i - d);
        int lmax = 2;
        while (delta(i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int s = lmax >> 1; s >= 1; s >>= 1) if (delta(i, i + (l + s) * d) > dmin) l += s;
        int j = i + l * d;
        int dn = delta(i, j);
        int s = 0, tt = l;
        do { tt = (tt + 1) >> 1; if (delta(i, i + (s + tt) * d) > dn) s += tt; } while (tt > 1);
        int gamma = i + s * d + std::min(d, 0);
        int lo = std::min(i, j), hi = std::max(i, j);
        t.left[i] = lo == gamma ? ~gamma : gamma;
        t.right[i] = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
        t.lastL[i] = gamma; t.lastR[i] = hi;
    }
    t.root = 0;
    refit(t, 0);
    return t;
}

static Tree implicitTree() {      // median split by index over the sorted leaves
    Tree t; int m = n - 1;
    t.left.reserve(m); t.right.reserve(m);
    struct J { int lo, hi, node; };
    std::vector<J> st;
    auto newNode = [&]() { t.left.push_back(0); t.right.push_back(0); t.lastL.push_back(0); t.lastR.push_back(0); t.boxL.push_back(Box()); t.boxR.push_back(Box()); return (int)t.left.size() - 1; };
    int root = newNode();
    st.push_back({ 0, n - 1, root });
    while (!st.empty()) {
        J j = st.back(); st.pop_back();
        int mid = (j.lo + j.hi) / 2;
        t.lastL[j.node] = mid; t.lastR[j.node] = j.hi;
        if (mid == j.lo) t.left[j.node] = ~j.lo; else { int c = newNode(); t.left[j.node] = c; st.push_back({ j.lo, mid, c }); }
        if (mid + 1 == j.hi) t.right[j.node] = ~j.hi; else { int c = newNode(); t.right[j.node] = c; st.push_back({ mid + 1, j.hi, c }); }
    }
    t.root = root;
    refit(t, root);
    return t;
}

// K > 0: a subtree of at most K leaves is not walked but tested leaf by leaf against the whole packet (one event of `size` leaves)
static void walkBrute(const Tree& t, int P, int K, const std::vector<int>& first, const char* name) {
    long long visits = 0, events = 0, tested = 0;
    int packets = 0;
    std::vector<int> st;
    auto sizeOf = [&](int node) { return std::max(t.lastL[node], t.lastR[node]) - first[node] + 1; };
    for (int base = 0; base < n; base += P, ++packets) {
        int cnt = std::min(P, n - base);
        st.clear(); st.push_back(t.root);
        while (!st.empty()) {
            int node = st.back(); st.pop_back();
            if (sizeOf(node) <= K) { ++events; tested += sizeOf(node); continue; }
            ++visits;
            bool anyL = false, anyR = false;
            for (int q = 0; q < cnt; ++q) {
                int i = base + q;
                anyL |= t.lastL[node] > i && ov(leaf[i], t.boxL[node]);
                anyR |= t.lastR[node] > i && ov(leaf[i], t.boxR[node]);
            }
            if (anyL && t.left[node] >= 0) st.push_back(t.left[node]);
            if (anyR && t.right[node] >= 0) st.push_back(t.right[node]);
        }
    }
    // instruction model from the SASS of k_lbvh_pairs: ~96 per node visit (x1.3 with two boxes per lane), ~30 + 28 per leaf for a direct test
    double perVisit = P == 64 ? 96 * 1.3 : 96, perLeaf = P == 64 ? 28 * 1.6 : 28;
    double instr = visits * perVisit + events * 30.0 + tested * perLeaf;
    printf("%-8s %2d/packet, direct below %2d leaves: %7.1f visits + %5.1f direct events (%6.1f leaves) per packet -> ~%.0f instructions per leaf\n", name, P, K,
           (double)visits / packets, (double)events / packets, (double)tested / packets, instr / n);
}

static void walk(const Tree& t, int P, const char* name) {
    long long visits = 0, pairs = 0, maxVisits = 0;
    int packets = 0;
    std::vector<int> st;
    for (int base = 0; base < n; base += P, ++packets) {
        int cnt = std::min(P, n - base);
        long long v = 0;
        st.clear(); st.push_back(t.root);
        while (!st.empty()) {
            int node = st.back(); st.pop_back();
            ++v;
            bool anyL = false, anyR = false;
            for (int q = 0; q < cnt; ++q) {
                int i = base + q;
                bool ol = t.lastL[node] > i && ov(leaf[i], t.boxL[node]);
                bool orr = t.lastR[node] > i && ov(leaf[i], t.boxR[node]);
                anyL |= ol; anyR |= orr;
                if (ol && t.left[node] < 0 && ~t.left[node] > i) ++pairs;
                if (orr && t.right[node] < 0 && ~t.right[node] > i) ++pairs;
            }
            if (anyL && t.left[node] >= 0) st.push_back(t.left[node]);
            if (anyR && t.right[node] >= 0) st.push_back(t.right[node]);
        }
        visits += v; maxVisits = std::max(maxVisits, v);
    }
    printf("%-10s %2d leaves/packet: %8.1f node visits per packet (max %lld), %6.2f per leaf, %lld pairs\n", name, P, (double)visits / packets, maxVisits,
           (double)visits / n, pairs);
}

int main(int argc, char** argv) {
    FILE* f = fopen(argc > 1 ? argv[1] : "/tmp/aabbs.bin", "rb");
    if (!f) { printf("no input\n"); return 1; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    n = (int)(sz / sizeof(Box));
    std::vector<Box> in(n);
    if (fread(in.data(), sizeof(Box), n, f) != (size_t)n) return 1;
    fclose(f);
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (auto& b : in) for (int k = 0; k < 3; ++k) { float c = (b.mn[k] + b.mx[k]) * 0.5f; lo[k] = std::min(lo[k], c); hi[k] = std::max(hi[k], c); }
    std::vector<uint32_t> k0(n);
    for (int i = 0; i < n; ++i) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) {
            float ext = hi[k] - lo[k], s = ext > 0.f ? 1024.f / ext : 0.f;
            float c = (in[i].mn[k] + in[i].mx[k]) * 0.5f;
            q[k] = (uint32_t)std::min(std::max((c - lo[k]) * s, 0.f), 1023.f);
        }
        k0[i] = (expandBits(q[0]) << 2) | (expandBits(q[1]) << 1) | expandBits(q[2]);
    }
    std::vector<int> order(n); std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return k0[a] < k0[b]; });
    leaf.resize(n); key.resize(n);
    for (int i = 0; i < n; ++i) { leaf[i] = in[order[i]]; key[i] = k0[order[i]]; }
    printf("%d leaves\n", n);
    Tree tk = karras();
    walk(tk, 32, "karras"); walk(tk, 64, "karras");
    {
        // first sorted position under every node (lastL / lastR hold the last ones)
        std::vector<int> first(tk.left.size());
        std::vector<int> order2; order2.reserve(tk.left.size());
        std::vector<int> stck{ tk.root };
        while (!stck.empty()) { int nd = stck.back(); stck.pop_back(); order2.push_back(nd); if (tk.left[nd] >= 0) stck.push_back(tk.left[nd]); if (tk.right[nd] >= 0) stck.push_back(tk.right[nd]); }
        for (int k = (int)order2.size() - 1; k >= 0; --k) { int nd = order2[k]; first[nd] = tk.left[nd] < 0 ? ~tk.left[nd] : first[tk.left[nd]]; }
        for (int P : { 32, 64 }) for (int K : { 0, 4, 8, 16, 32 }) walkBrute(tk, P, K, first, "karras");
    }
    Tree ti = implicitTree();
    walk(ti, 32, "implicit"); walk(ti, 64, "implicit");
    return 0;
}
