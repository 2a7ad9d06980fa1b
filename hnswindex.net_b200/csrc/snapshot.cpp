// snapshot.cpp -- streaming reader / writer for the HNSWIndex.Net snapshot file (see snapshot.hpp).
#include "snapshot.hpp"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace hnsw {
namespace {

inline int vsize(uint64_t v) {
    int n = 1;
    while (v >= 0x80) { v >>= 7; ++n; }
    return n;
}
inline uint64_t sx(int32_t v) { return (uint64_t)(int64_t)v; }  // int32 -> sign-extended varint payload

struct FileSink {
    FILE* f;
    std::vector<uint8_t> buf;
    explicit FileSink(FILE* fp) : f(fp) { buf.reserve(1 << 20); }
    void flush() {
        if (!buf.empty() && std::fwrite(buf.data(), 1, buf.size(), f) != buf.size()) throw std::runtime_error("snapshot: write failed");
        buf.clear();
    }
    void byte(uint8_t b) {
        buf.push_back(b);
        if (buf.size() >= (1 << 20)) flush();
    }
    void put(const void* p, size_t n) {
        if (buf.size() + n >= (1 << 20)) flush();
        if (n >= (1 << 20)) {
            if (std::fwrite(p, 1, n, f) != n) throw std::runtime_error("snapshot: write failed");
        } else {
            const uint8_t* b = (const uint8_t*)p;
            buf.insert(buf.end(), b, b + n);
        }
    }
    void varint(uint64_t v) {
        while (v >= 0x80) { byte((uint8_t)((v & 0x7f) | 0x80)); v >>= 7; }
        byte((uint8_t)v);
    }
    void tag(int field, int wt) { varint(((uint64_t)field << 3) | (uint64_t)wt); }
};

// ---- sizes -------------------------------------------------------------------
uint64_t int32s_size(int field_tag_bytes, const int* v, int n, bool packed) {
    if (n == 0) return 0;
    uint64_t payload = 0;
    for (int i = 0; i < n; i++) payload += vsize(sx(v[i]));
    if (packed) return field_tag_bytes + vsize(payload) + payload;
    return payload + (uint64_t)field_tag_bytes * n;
}
uint64_t edge_list_body(const int* ids, int n, bool packed) {
    uint64_t b = int32s_size(1, ids, n, packed);
    if (n != 0) b += 1 + vsize((uint64_t)n);
    return b;
}
uint64_t item_body(int dim, bool packed) {
    if (dim == 0) return 0;
    if (packed) return 1 + vsize((uint64_t)dim * 4) + (uint64_t)dim * 4;
    return (uint64_t)dim * 5;
}

// ---- emit ----------------------------------------------------------------------
void emit_int32s(FileSink& s, int field, const int* v, int n, bool packed) {
    if (n == 0) return;
    if (packed) {
        uint64_t payload = 0;
        for (int i = 0; i < n; i++) payload += vsize(sx(v[i]));
        s.tag(field, 2);
        s.varint(payload);
        for (int i = 0; i < n; i++) s.varint(sx(v[i]));
    } else {
        for (int i = 0; i < n; i++) { s.tag(field, 0); s.varint(sx(v[i])); }
    }
}
void emit_edge_list(FileSink& s, int field, const int* ids, int n, bool packed) {
    s.tag(field, 2);
    s.varint(edge_list_body(ids, n, packed));
    emit_int32s(s, 1, ids, n, packed);
    if (n != 0) { s.tag(2, 0); s.varint((uint64_t)n); }
}
inline int row_count(const int* row, int stride) {
    int n = 0;
    while (n < stride && row[n] >= 0) ++n;
    return n;
}

struct InEdges {  // transposed adjacency of one layer (sources ascending)
    std::vector<uint32_t> off;
    std::vector<int> ids;
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    Reader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        for (;;) {
            if (p >= end) throw std::runtime_error("snapshot: truncated varint");
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
            if (shift > 63) throw std::runtime_error("snapshot: malformed varint");
        }
        return v;
    }
    uint32_t fixed32() {
        if (end - p < 4) throw std::runtime_error("snapshot: truncated fixed32");
        uint32_t v; std::memcpy(&v, p, 4); p += 4; return v;
    }
    uint64_t fixed64() {
        if (end - p < 8) throw std::runtime_error("snapshot: truncated fixed64");
        uint64_t v; std::memcpy(&v, p, 8); p += 8; return v;
    }
    Reader sub() {
        const uint64_t n = varint();
        if ((uint64_t)(end - p) < n) throw std::runtime_error("snapshot: truncated sub-message");
        Reader r(p, p + n);
        p += n;
        return r;
    }
    void skip(int wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: fixed64(); break;
            case 2: sub(); break;
            case 5: fixed32(); break;
            default: throw std::runtime_error("snapshot: unsupported wire type");
        }
    }
    bool next(int& field, int& wt) {
        if (done()) return false;
        const uint64_t t = varint();
        field = (int)(t >> 3);
        wt = (int)(t & 7);
        return true;
    }
};

void read_int32s(Reader& r, int wt, std::vector<int>& out) {
    if (wt == 0) out.push_back((int32_t)r.varint());
    else if (wt == 2) { Reader s = r.sub(); while (!s.done()) out.push_back((int32_t)s.varint()); }
    else r.skip(wt);
}

}  // namespace

void write_snapshot(const HostGraph& g, const std::string& path, bool packed) {
    const int n = g.length;
    // levels above 0 present?
    int top = 0;
    for (int i = 0; i < n; i++) top = std::max(top, g.levels[i]);

    // in-edges (Node.cs:24-25) are the transpose of the out-edges; only written when AllowRemovals
    std::vector<InEdges> in(g.params.AllowRemovals ? top + 1 : 0);
    for (size_t l = 0; l < in.size(); l++) {
        InEdges& ie = in[l];
        ie.off.assign((size_t)n + 1, 0);
        auto row_of = [&](int i, int& stride) -> const int* {
            if (g.levels[i] < (int)l) { stride = 0; return nullptr; }
            if (l == 0) { stride = g.stride0; return g.adj0.data() + (size_t)i * g.stride0; }
            stride = g.strideU;
            return g.adjU.data() + ((size_t)g.upper_off[i] + (l - 1)) * g.strideU;
        };
        for (int i = 0; i < n; i++) {
            int stride; const int* row = row_of(i, stride);
            for (int j = 0; j < stride && row[j] >= 0; j++) ie.off[(size_t)row[j] + 1]++;
        }
        for (int i = 0; i < n; i++) ie.off[i + 1] += ie.off[i];
        ie.ids.resize(ie.off[n]);
        std::vector<uint32_t> fill(ie.off.begin(), ie.off.end() - 1);
        for (int i = 0; i < n; i++) {
            int stride; const int* row = row_of(i, stride);
            for (int j = 0; j < stride && row[j] >= 0; j++) ie.ids[fill[row[j]]++] = i;
        }
    }

    auto out_row = [&](int i, int l, int& cnt) -> const int* {
        const int* row;
        int stride;
        if (l == 0) { stride = g.stride0; row = g.adj0.data() + (size_t)i * g.stride0; }
        else { stride = g.strideU; row = g.adjU.data() + ((size_t)g.upper_off[i] + (l - 1)) * g.strideU; }
        cnt = row_count(row, stride);
        return row;
    };
    const bool any_removed = !g.removed.empty();
    auto is_removed = [&](int i) { return any_removed && g.removed[i] != 0; };
    auto node_body = [&](int i) -> uint64_t {
        uint64_t b = 0;
        if (i != 0) b += 1 + vsize(sx(i));
        if (is_removed(i)) b += 2;  // IsRemoved = true (Node.cs:18-19)
        for (int l = 0; l <= g.levels[i]; l++) {
            int cnt; const int* row = out_row(i, l, cnt);
            const uint64_t eb = edge_list_body(row, cnt, packed);
            b += 1 + vsize(eb) + eb;
        }
        if (g.params.AllowRemovals)
            for (int l = 0; l <= g.levels[i]; l++) {
                const InEdges& ie = in[l];
                const uint64_t eb = edge_list_body(ie.ids.data() + ie.off[i], (int)(ie.off[i + 1] - ie.off[i]), packed);
                b += 1 + vsize(eb) + eb;
            }
        return b;
    };

    // ---- sizes of the two top-level messages
    const Parameters& p = g.params;
    uint64_t params_body = 0;
    if (p.MaxEdges) params_body += 1 + vsize(sx(p.MaxEdges));
    if (p.DistributionRate != 0.0) params_body += 9;
    if (p.MinNN) params_body += 1 + vsize(sx(p.MinNN));
    if (p.MaxCandidates) params_body += 1 + vsize(sx(p.MaxCandidates));
    if (p.RemoveMaxCandidates) params_body += 1 + vsize(sx(p.RemoveMaxCandidates));
    if (p.CollectionSize) params_body += 1 + vsize(sx(p.CollectionSize));
    if (p.RandomSeed) params_body += 1 + vsize(sx(p.RandomSeed));
    if (p.AllowRemovals) params_body += 2;

    // ActiveSet.dense[..count]: insertion order, a removal moves the last entry into the hole (ActiveSet.cs:79-93)
    std::vector<int> active;
    if (!g.active.empty() || any_removed) active = g.active;
    else { active.resize(n); for (int i = 0; i < n; i++) active[i] = i; }
    const int nactive = (int)active.size();
    // GraphData.RemovedIndexes: a ConcurrentStack enumerates from its top [inferred: protobuf-net writes in enumeration order]
    std::vector<int> removed_top_first(g.removed_stack.rbegin(), g.removed_stack.rend());

    uint64_t data_body = 0;
    for (int i = 0; i < n; i++) { const uint64_t nb = node_body(i); data_body += 1 + vsize(nb) + nb; }
    data_body += int32s_size(1, active.data(), nactive, packed);
    data_body += int32s_size(1, removed_top_first.data(), (int)removed_top_first.size(), packed);
    const uint64_t ib = item_body(g.dim, packed);
    data_body += (uint64_t)n * (1 + vsize(ib) + ib);
    if (g.entry != 0) data_body += 1 + vsize(sx(g.entry));
    if (g.capacity) data_body += 1 + vsize(sx(g.capacity));
    if (g.length) data_body += 1 + vsize(sx(g.length));
    if (nactive) data_body += 1 + vsize(sx(nactive));

    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("snapshot: cannot create " + path);
    try {
        FileSink s(f);
        s.tag(1, 2);
        s.varint(params_body);
        if (p.MaxEdges) { s.tag(1, 0); s.varint(sx(p.MaxEdges)); }
        if (p.DistributionRate != 0.0) { s.tag(2, 1); s.put(&p.DistributionRate, 8); }
        if (p.MinNN) { s.tag(3, 0); s.varint(sx(p.MinNN)); }
        if (p.MaxCandidates) { s.tag(4, 0); s.varint(sx(p.MaxCandidates)); }
        if (p.RemoveMaxCandidates) { s.tag(5, 0); s.varint(sx(p.RemoveMaxCandidates)); }
        if (p.CollectionSize) { s.tag(6, 0); s.varint(sx(p.CollectionSize)); }
        if (p.RandomSeed) { s.tag(7, 0); s.varint(sx(p.RandomSeed)); }
        if (p.AllowRemovals) { s.tag(8, 0); s.varint(1); }

        s.tag(2, 2);
        s.varint(data_body);
        for (int i = 0; i < n; i++) {
            s.tag(1, 2);
            s.varint(node_body(i));
            if (i != 0) { s.tag(1, 0); s.varint(sx(i)); }
            if (is_removed(i)) { s.tag(2, 0); s.varint(1); }
            for (int l = 0; l <= g.levels[i]; l++) {
                int cnt; const int* row = out_row(i, l, cnt);
                emit_edge_list(s, 3, row, cnt, packed);
            }
            if (p.AllowRemovals)
                for (int l = 0; l <= g.levels[i]; l++) {
                    const InEdges& ie = in[l];
                    emit_edge_list(s, 4, ie.ids.data() + ie.off[i], (int)(ie.off[i + 1] - ie.off[i]), packed);
                }
        }
        emit_int32s(s, 2, active.data(), nactive, packed);
        for (int i = 0; i < n; i++) {
            const float* v = g.vectors.data() + (size_t)i * g.dim;
            s.tag(3, 2);
            s.varint(ib);
            if (g.dim == 0) continue;
            if (packed) {
                s.tag(1, 2);
                s.varint((uint64_t)g.dim * 4);
                s.put(v, (size_t)g.dim * 4);
            } else {
                for (int d = 0; d < g.dim; d++) { s.byte(0x0D); s.put(v + d, 4); }
            }
        }
        emit_int32s(s, 4, removed_top_first.data(), (int)removed_top_first.size(), packed);
        if (g.entry != 0) { s.tag(5, 0); s.varint(sx(g.entry)); }
        if (g.capacity) { s.tag(6, 0); s.varint(sx(g.capacity)); }
        if (g.length) { s.tag(7, 0); s.varint(sx(g.length)); }
        if (nactive) { s.tag(8, 0); s.varint(sx(nactive)); }
        s.flush();
    } catch (...) {
        std::fclose(f);
        throw;
    }
    if (std::fclose(f) != 0) throw std::runtime_error("snapshot: close failed");
}

void read_snapshot(const std::string& path, HostGraph& g) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("snapshot: cannot open " + path);
    std::vector<uint8_t> bytes;
    {
        std::fseek(f, 0, SEEK_END);
        const long long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        bytes.resize((size_t)(sz < 0 ? 0 : sz));
        const size_t got = bytes.empty() ? 0 : std::fread(bytes.data(), 1, bytes.size(), f);
        std::fclose(f);
        if (got != bytes.size()) throw std::runtime_error("snapshot: short read");
    }
    g = HostGraph();
    Parameters& p = g.params;
    bool have_params = false, have_data = false, have_allow = false, have_entry = false, any_in = false;

    // pass 1: parameters (strides depend on MaxEdges)
    {
        Reader root(bytes.data(), bytes.data() + bytes.size());
        int field, wt;
        while (root.next(field, wt)) {
            if (field == 1 && wt == 2) {
                have_params = true;
                Reader r = root.sub();
                int f2, w2;
                while (r.next(f2, w2)) {
                    if (f2 == 1 && w2 == 0) p.MaxEdges = (int32_t)r.varint();
                    else if (f2 == 2 && w2 == 1) { const uint64_t b = r.fixed64(); std::memcpy(&p.DistributionRate, &b, 8); }
                    else if (f2 == 3 && w2 == 0) p.MinNN = (int32_t)r.varint();
                    else if (f2 == 4 && w2 == 0) p.MaxCandidates = (int32_t)r.varint();
                    else if (f2 == 5 && w2 == 0) p.RemoveMaxCandidates = (int32_t)r.varint();
                    else if (f2 == 6 && w2 == 0) p.CollectionSize = (int32_t)r.varint();
                    else if (f2 == 7 && w2 == 0) p.RandomSeed = (int32_t)r.varint();
                    else if (f2 == 8 && w2 == 0) { p.AllowRemovals = r.varint() != 0; have_allow = true; }
                    else r.skip(w2);
                }
            } else root.skip(wt);
        }
    }
    if (!have_params) throw std::runtime_error("Parameters cannot be null during deserialization.");  // HNSWIndex.cs:37-38
    if (p.MaxEdges < 1) throw std::runtime_error("snapshot: MaxEdges must be positive");
    g.stride0 = stride_for(2 * p.MaxEdges);
    g.strideU = stride_for(p.MaxEdges);

    int count_field = 0;
    std::vector<int> row, active, removed_top_first;
    std::vector<int> removed_nodes;
    {
        Reader root(bytes.data(), bytes.data() + bytes.size());
        int field, wt;
        while (root.next(field, wt)) {
            if (!(field == 2 && wt == 2)) { root.skip(wt); continue; }
            have_data = true;
            Reader r = root.sub();
            int f2, w2;
            int items_seen = 0;
            while (r.next(f2, w2)) {
                if (f2 == 1 && w2 == 2) {  // Node
                    Reader nr = r.sub();
                    const int idx = (int)g.levels.size();
                    int id = 0, f3, w3, nout = 0;
                    bool removed = false;
                    g.adj0.resize((size_t)(idx + 1) * g.stride0, -1);
                    g.upper_off.push_back(-1);
                    while (nr.next(f3, w3)) {
                        if (f3 == 1 && w3 == 0) id = (int32_t)nr.varint();
                        else if (f3 == 2 && w3 == 0) removed = nr.varint() != 0;
                        else if (f3 == 3 && w3 == 2) {
                            Reader er = nr.sub();
                            row.clear();
                            int cnt = 0, f4, w4;
                            while (er.next(f4, w4)) {
                                if (f4 == 1) read_int32s(er, w4, row);
                                else if (f4 == 2 && w4 == 0) cnt = (int32_t)er.varint();
                                else er.skip(w4);
                            }
                            if (cnt < 0 || cnt > (int)row.size()) throw std::runtime_error("snapshot: EdgeList.Count exceeds its Buffer");
                            const int stride = nout == 0 ? g.stride0 : g.strideU;
                            const int cap = nout == 0 ? 2 * p.MaxEdges : p.MaxEdges;
                            if (cnt > cap) throw std::runtime_error("snapshot: a row holds more than MaxEdges(layer) links");
                            int* dst;
                            if (nout == 0) dst = g.adj0.data() + (size_t)idx * g.stride0;
                            else {
                                if (nout == 1) g.upper_off[idx] = (int)(g.adjU.size() / g.strideU);
                                g.adjU.resize(g.adjU.size() + g.strideU, -1);
                                dst = g.adjU.data() + g.adjU.size() - g.strideU;
                            }
                            for (int j = 0; j < cnt; j++) dst[j] = row[j];  // truncate Buffer to Count
                            (void)stride;
                            nout++;
                        } else if (f3 == 4 && w3 == 2) { any_in = true; nr.skip(w3); }
                        else nr.skip(w3);
                    }
                    if (removed) removed_nodes.push_back(idx);
                    if (id != idx) throw std::runtime_error("snapshot: node ids must equal their slot");
                    if (nout < 1) throw std::runtime_error("snapshot: node without layer-0 edge list");
                    g.levels.push_back(nout - 1);
                } else if (f2 == 2) read_int32s(r, w2, active);
                else if (f2 == 3 && w2 == 2) {  // item
                    Reader ir = r.sub();
                    const size_t before = g.vectors.size();
                    int f3, w3;
                    while (ir.next(f3, w3)) {
                        if (f3 == 1 && w3 == 5) { const uint32_t b = ir.fixed32(); float v; std::memcpy(&v, &b, 4); g.vectors.push_back(v); }
                        else if (f3 == 1 && w3 == 2) {
                            Reader s = ir.sub();
                            const size_t nb = (size_t)(s.end - s.p);
                            if (nb % 4) throw std::runtime_error("snapshot: packed float payload not a multiple of 4");
                            const size_t old = g.vectors.size();
                            g.vectors.resize(old + nb / 4);
                            std::memcpy(g.vectors.data() + old, s.p, nb);
                        } else ir.skip(w3);
                    }
                    const int d = (int)(g.vectors.size() - before);
                    if (items_seen == 0) g.dim = d;
                    else if (d != g.dim) throw std::runtime_error("snapshot: items have different lengths");
                    items_seen++;
                } else if (f2 == 4) read_int32s(r, w2, removed_top_first);
                else if (f2 == 5 && w2 == 0) { g.entry = (int32_t)r.varint(); have_entry = true; }
                else if (f2 == 6 && w2 == 0) g.capacity = (int32_t)r.varint();
                else if (f2 == 7 && w2 == 0) g.length = (int32_t)r.varint();
                else if (f2 == 8 && w2 == 0) count_field = (int32_t)r.varint();
                else r.skip(w2);
            }
            if (items_seen != (int)g.levels.size()) throw std::runtime_error("snapshot: Nodes and Items differ in length");
        }
    }
    if (!have_data) throw std::runtime_error("Data cannot be null during deserialization.");  // HNSWIndex.cs:40-41
    const int n = (int)g.levels.size();
    // Latent reference quirks handled deliberately: EntryPointId == 0 is omitted
    // on write (zero default), and so is AllowRemovals == false.
    if (!have_entry) g.entry = n > 0 ? 0 : -1;
    if (!have_allow) p.AllowRemovals = any_in || n == 0;
    if (g.length == 0) g.length = n;
    if (g.length != n) throw std::runtime_error("snapshot: Length does not match the number of nodes");
    // removal state: the reference leaves the rows of a removed node stale (nothing points at it); they are emptied here
    if (!removed_nodes.empty() || !removed_top_first.empty()) {
        g.removed.assign((size_t)n, 0);
        for (int i : removed_nodes) {
            g.removed[i] = 1;
            for (int j = 0; j < g.stride0; j++) g.adj0[(size_t)i * g.stride0 + j] = -1;
            for (int l = 1; l <= g.levels[i]; l++)
                for (int j = 0; j < g.strideU; j++) g.adjU[((size_t)g.upper_off[i] + (l - 1)) * g.strideU + j] = -1;
        }
        g.removed_stack.assign(removed_top_first.rbegin(), removed_top_first.rend());
        g.active = active;
        if ((int)active.size() + (int)removed_nodes.size() != n) throw std::runtime_error("snapshot: ActiveNodes and IsRemoved flags disagree");
    } else if ((int)active.size() == n) {
        bool identity = true;
        for (int i = 0; identity && i < n; i++) identity = active[i] == i;
        if (!identity) g.active = active;
    }
    const int nactive = n - (int)removed_nodes.size();
    if (count_field != 0 && count_field != nactive) throw std::runtime_error("snapshot: Count does not match the number of active nodes");
    if (g.capacity < n) g.capacity = std::max(n, p.CollectionSize);
    if (g.capacity > (1 << 30)) throw std::runtime_error("snapshot: Capacity is not a plausible collection size");
    // EntryPointId: -1 (empty index) or a live slot -- checked whether or not any node is active: the engine indexes its
    // per-node arrays with it
    if (g.entry < -1 || g.entry >= n) throw std::runtime_error("snapshot: EntryPointId out of range");
    if (nactive > 0 && g.entry < 0) throw std::runtime_error("snapshot: EntryPointId out of range");
    if (g.entry >= 0 && !g.removed.empty() && g.removed[g.entry]) throw std::runtime_error("snapshot: the entry point is a removed node");
    // ActiveNodes: live slots, each once
    if (!g.active.empty()) {
        std::vector<unsigned char> seen((size_t)n, 0);
        for (int id : g.active) {
            if (id < 0 || id >= n || seen[id] || (!g.removed.empty() && g.removed[id])) throw std::runtime_error("snapshot: inconsistent ActiveNodes");
            seen[id] = 1;
        }
    }
    // RemovedIndexes is the free-slot stack Add pops from (GraphData.cs:84-91): every entry must be a removed slot of this
    // file, each exactly once, and every removed slot must be on it -- otherwise a later Add would write outside the arrays
    // or hand one slot to two items
    if (g.removed_stack.size() != removed_nodes.size()) throw std::runtime_error("snapshot: RemovedIndexes and IsRemoved flags disagree");
    {
        std::vector<unsigned char> seen((size_t)n, 0);
        for (int id : g.removed_stack) {
            if (id < 0 || id >= n || !g.removed[id] || seen[id]) throw std::runtime_error("snapshot: inconsistent RemovedIndexes");
            seen[id] = 1;
        }
    }
    for (int v : g.adj0) if (v >= n) throw std::runtime_error("snapshot: edge to a node that does not exist");
    for (int v : g.adjU) if (v >= n) throw std::runtime_error("snapshot: edge to a node that does not exist");
}

}  // namespace hnsw
