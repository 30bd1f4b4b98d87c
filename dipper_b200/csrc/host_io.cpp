// Host-side pieces of dipper_b200 (include/dipper_host.h): encoders, Newick writer /
// backbone reader.  Plain C++17; no CUDA here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/dipper_b200.h"
#include "../../include/dipper_host.h"

namespace dipb {
void set_error(const char* fmt, ...);
}

namespace {

struct Code4 {
    unsigned char t[256];
    Code4() {
        for (int i = 0; i < 256; i++) t[i] = 4;
        t['A'] = 0; t['C'] = 1; t['G'] = 2; t['T'] = 3; t['U'] = 3;
    }
};
const Code4 kCode4;

void append_g(std::string& s, double v) {
    char buf[64];
    int n = snprintf(buf, sizeof buf, "%g", v);  // ostream<<double default formatting
    s.append(buf, (size_t)n);
}

char* dup_string(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

}  // namespace

extern "C" {

void dipb_pack4(const char* seq, size_t len, uint64_t* out) {
    size_t nw = (len + 15) / 16;
    for (size_t w = 0; w < nw; w++) {
        uint64_t v = 0;
        size_t end = w * 16 + 16 < len ? w * 16 + 16 : len;
        for (size_t s = w * 16; s < end; s++) v |= (uint64_t)kCode4.t[(unsigned char)seq[s]] << (4 * (s & 15));
        out[w] = v;
    }
}

void dipb_pack2(const char* seq, size_t len, uint64_t* out) {
    size_t nw = (len + 31) / 32;
    for (size_t w = 0; w < nw; w++) {
        uint64_t v = 0;
        size_t end = w * 32 + 32 < len ? w * 32 + 32 : len;
        for (size_t s = w * 32; s < end; s++) {
            uint64_t c = kCode4.t[(unsigned char)seq[s]];
            v |= (c & 3 & (uint64_t)-(int64_t)(c < 4)) << (2 * (s & 31));
        }
        out[w] = v;
    }
}

char* dipb_nj_newick(int n, const int32_t* c0, const int32_t* c1, const double* l0, const double* l1,
                     const char* const* names) {
    std::string s;
    s.reserve((size_t)n * 24);
    struct Frame { int node, stage; };
    std::vector<Frame> st;
    st.push_back({2 * n - 2, 0});
    while (!st.empty()) {
        Frame& f = st.back();
        if (f.node < n) { s += names[f.node]; st.pop_back(); continue; }
        int k = f.node - n;
        if (f.stage == 0) { f.stage = 1; s += '('; st.push_back({c0[k], 0}); }
        else if (f.stage == 1) { f.stage = 2; s += ':'; append_g(s, l0[k]); s += ','; st.push_back({c1[k], 0}); }
        else { s += ':'; append_g(s, l1[k]); s += ')'; st.pop_back(); }
    }
    s += ";\n";
    return dup_string(s);
}

char* dipb_tree_newick(int n_nodes, int root_node, const int32_t* head, const int32_t* e, const int32_t* nxt,
                       const double* len, const char* const* names) {
    (void)n_nodes;
    std::string s;
    struct Frame { int node, from, slot; };
    std::vector<Frame> st;
    st.push_back({root_node, -1, -2});
    while (!st.empty()) {
        Frame& f = st.back();
        if (nxt[head[f.node]] == -1) { s += names[f.node]; st.pop_back(); continue; }
        int q;
        bool fresh = f.slot == -2;
        if (fresh) { s += '('; q = head[f.node]; }
        else { s += ':'; append_g(s, len[f.slot]); q = nxt[f.slot]; }
        while (q != -1 && e[q] == f.from) q = nxt[q];
        if (q == -1) { s += ')'; st.pop_back(); continue; }
        if (!fresh) s += ',';
        f.slot = q;
        int child = e[q], me = f.node;
        st.push_back({child, me, -2});
    }
    s += ";\n";
    return dup_string(s);
}

void dipb_free_str(char* s) { free(s); }

int dipb_backbone_from_newick(const char* newick, int total_leaves, int32_t* head, int32_t* e, int32_t* nxt,
                              int32_t* belong, double* len, char** leaf_names_out) {
    if (!newick || total_leaves < 2 || !head || !e || !nxt || !belong || !len) {
        dipb::set_error("dipb_backbone_from_newick: bad argument");
        return DIPB_E_ARG;
    }
    struct N { int parent; int idx; double bl; std::vector<int> ch; std::string name; };
    std::vector<N> nodes;
    std::vector<int> stack;
    int next_internal = total_leaves, next_leaf = 0;
    int cur = -1;  // node whose label/length is being read
    const char* p = newick;
    int root = -1;
    while (*p && *p != ';') {
        char c = *p;
        if (c == '(') {
            N nd; nd.parent = stack.empty() ? -1 : stack.back(); nd.idx = next_internal++; nd.bl = 0;
            nodes.push_back(nd);
            int id = (int)nodes.size() - 1;
            if (nd.parent >= 0) nodes[nd.parent].ch.push_back(id); else root = id;
            stack.push_back(id);
            cur = -1; p++;
        } else if (c == ',') { cur = -1; p++; }
        else if (c == ')') {
            if (stack.empty()) { dipb::set_error("newick: unbalanced ')'"); return DIPB_E_ARG; }
            cur = stack.back(); stack.pop_back(); p++;
        } else if (c == ':') {
            p++;
            char* endp = nullptr;
            float f = strtof(p, &endp);  // src/tree.cpp:268,289 parse with stof
            if (endp == p) { dipb::set_error("newick: missing branch length"); return DIPB_E_ARG; }
            if (cur >= 0) nodes[cur].bl = (double)f;
            p = endp;
        } else if (c == ' ' || c == '\n' || c == '\r' || c == '\t') { p++; }
        else {
            std::string label;
            if (c == '\'') { p++; while (*p && *p != '\'') label += *p++; if (*p == '\'') p++; }
            else while (*p && !strchr(":,();", *p)) label += *p++;
            if (cur == -1) {
                if (stack.empty()) { dipb::set_error("newick: leaf outside parentheses"); return DIPB_E_ARG; }
                N nd; nd.parent = stack.back(); nd.idx = next_leaf++; nd.bl = 0; nd.name = label;
                nodes.push_back(nd);
                cur = (int)nodes.size() - 1;
                nodes[nd.parent].ch.push_back(cur);
            }  // labels on internal nodes are ignored, as in the reference
        }
    }
    if (!stack.empty() || root < 0) { dipb::set_error("newick: unbalanced parentheses"); return DIPB_E_ARG; }
    const int B = next_leaf;
    if (B > total_leaves || 2 * (size_t)(nodes.size() - 1) > 8 * (size_t)total_leaves) {
        dipb::set_error("newick: %d leaves do not fit total_leaves=%d", B, total_leaves);
        return DIPB_E_ARG;
    }
    for (int i = 0; i < 2 * total_leaves; i++) head[i] = -1;
    for (int i = 0; i < 8 * total_leaves; i++) { e[i] = -1; nxt[i] = -1; belong[i] = -1; len[i] = 2; }
    // post-order, two slots per non-root node: child->parent then parent->child
    struct F { int node; size_t k; };
    std::vector<F> st;
    st.push_back({root, 0});
    int edge = 0;
    while (!st.empty()) {
        F& f = st.back();
        if (f.k < nodes[f.node].ch.size()) { int c2 = nodes[f.node].ch[f.k++]; st.push_back({c2, 0}); continue; }
        int v = f.node;
        st.pop_back();
        if (nodes[v].parent < 0) continue;
        int x = nodes[v].idx, y = nodes[nodes[v].parent].idx;
        e[edge] = y; len[edge] = nodes[v].bl; belong[edge] = x; nxt[edge] = head[x]; head[x] = edge; edge++;
        e[edge] = x; len[edge] = nodes[v].bl; belong[edge] = y; nxt[edge] = head[y]; head[y] = edge; edge++;
    }
    if (leaf_names_out) {
        std::vector<std::string> nm(B);
        for (auto& nd : nodes) if (nd.ch.empty() && nd.idx < B) nm[nd.idx] = nd.name;
        std::string all;
        for (int i = 0; i < B; i++) { all += nm[i]; all += '\n'; }
        *leaf_names_out = dup_string(all);
    }
    return B;
}

}  // extern "C"
