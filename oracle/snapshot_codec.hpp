// =============================================================================
// oracle/snapshot_codec.hpp -- TEST INFRASTRUCTURE ONLY (see hnsw_oracle.hpp).
//
// Reader/writer for the file HNSWIndex.Serialize produces (HNSWIndex.cs:210-229):
// a protobuf message written by protobuf-net 3.2.52 (NuGet dependency,
// HNSWIndex.csproj:37 -- NOT vendored in /root/reference) from the contracts in
// HNSWIndexSnapshot.cs:12-16, GraphDataSnapshot.cs:13-35, Node.cs:9-36,
// HNSWParameters.cs:12-55 and NestedListWrapper.cs:14-20.
//
// The wire format itself is the published protobuf encoding.  What protobuf-net
// chooses inside it (zero-default omission, unpacked repeated scalars) is
// restated from its documentation; no .NET-written file exists here, so the
// byte layout is UNPINNED.  The reader therefore accepts packed and unpacked
// repeated scalars and absent zero-valued fields.
// =============================================================================
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "hnsw_oracle.hpp"

namespace hnsw_oracle {
namespace pb {

struct Writer {
    std::string buf;
    void varint(uint64_t v) {
        while (v >= 0x80) { buf.push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
        buf.push_back((char)v);
    }
    void tag(int field, int wt) { varint(((uint64_t)field << 3) | (uint64_t)wt); }
    void int32_field(int field, int32_t v) { tag(field, 0); varint((uint64_t)(int64_t)v); }  // negative -> 10 bytes
    void int32_field_nz(int field, int32_t v) { if (v != 0) int32_field(field, v); }
    void bool_field_nz(int field, bool v) { if (v) { tag(field, 0); varint(1); } }
    void double_field_nz(int field, double v) {
        if (v == 0.0) return;
        tag(field, 1);
        uint64_t bits; std::memcpy(&bits, &v, 8);
        for (int i = 0; i < 8; i++) buf.push_back((char)((bits >> (8 * i)) & 0xff));
    }
    void bytes_field(int field, const std::string& s) { tag(field, 2); varint(s.size()); buf.append(s); }
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    Reader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0; int shift = 0;
        for (;;) {
            if (p >= end) throw std::runtime_error("snapshot: truncated varint");
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
            if (shift > 63) throw std::runtime_error("snapshot: varint too long");
        }
        return v;
    }
    uint32_t fixed32() { if (end - p < 4) throw std::runtime_error("snapshot: truncated"); uint32_t v; std::memcpy(&v, p, 4); p += 4; return v; }
    uint64_t fixed64() { if (end - p < 8) throw std::runtime_error("snapshot: truncated"); uint64_t v; std::memcpy(&v, p, 8); p += 8; return v; }
    Reader sub() {
        uint64_t n = varint();
        if ((uint64_t)(end - p) < n) throw std::runtime_error("snapshot: truncated sub-message");
        Reader r(p, p + n); p += n; return r;
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
};

inline void read_int32s(Reader& r, int wt, std::vector<int>& out) {
    if (wt == 0) out.push_back((int32_t)r.varint());
    else if (wt == 2) { Reader s = r.sub(); while (!s.done()) out.push_back((int32_t)s.varint()); }
    else r.skip(wt);
}
inline void read_floats(Reader& r, int wt, std::vector<float>& out) {
    if (wt == 5) { uint32_t b = r.fixed32(); float f; std::memcpy(&f, &b, 4); out.push_back(f); }
    else if (wt == 2) { Reader s = r.sub(); while (!s.done()) { uint32_t b = s.fixed32(); float f; std::memcpy(&f, &b, 4); out.push_back(f); } }
    else r.skip(wt);
}

}  // namespace pb

// ---- writer -----------------------------------------------------------------
inline void write_edge_list(pb::Writer& w, int field, const EdgeList& e, bool packed) {
    pb::Writer s;
    if (packed) {
        if (!e.Buffer.empty()) {
            pb::Writer b;
            for (int v : e.Buffer) b.varint((uint64_t)(int64_t)v);
            s.bytes_field(1, b.buf);
        }
    } else {
        for (int v : e.Buffer) s.int32_field(1, v);  // whole backing array, incl. stale slots (Node.cs:33-34)
    }
    s.int32_field_nz(2, e.Count);
    w.bytes_field(field, s.buf);
}

inline void serialize(const Index& ix, const std::string& path, bool packed = false) {
    pb::Writer params;
    params.int32_field_nz(1, ix.p.MaxEdges);
    params.double_field_nz(2, ix.p.DistributionRate);
    params.int32_field_nz(3, ix.p.MinNN);
    params.int32_field_nz(4, ix.p.MaxCandidates);
    params.int32_field_nz(5, ix.p.RemoveMaxCandidates);
    params.int32_field_nz(6, ix.p.CollectionSize);
    params.int32_field_nz(7, ix.p.RandomSeed);
    params.bool_field_nz(8, ix.p.AllowRemovals);

    pb::Writer data;
    for (auto& np : ix.Nodes) {  // GraphDataSnapshot.cs:61 (non-null nodes in slot order)
        pb::Writer n;
        n.int32_field_nz(1, np->Id);
        n.bool_field_nz(2, np->IsRemoved);
        for (auto& e : np->OutEdges) write_edge_list(n, 3, e, packed);
        for (auto& e : np->InEdges) write_edge_list(n, 4, e, packed);
        data.bytes_field(1, n.buf);
    }
    if (packed) {
        pb::Writer a; for (int id : ix.ActiveIds) a.varint((uint64_t)(int64_t)id);
        if (!ix.ActiveIds.empty()) data.bytes_field(2, a.buf);
    } else {
        for (int id : ix.ActiveIds) data.int32_field(2, id);
    }
    for (int i = 0; i < ix.Length; i++) {  // GraphDataSnapshot.cs:62 (one wrapper per item)
        pb::Writer it;
        const float* v = ix.item(i);
        if (packed) {
            it.tag(1, 2); it.varint((uint64_t)ix.dim * 4);
            it.buf.append((const char*)v, (size_t)ix.dim * 4);
        } else {
            for (int d = 0; d < ix.dim; d++) { it.tag(1, 5); it.buf.append((const char*)(v + d), 4); }
        }
        data.bytes_field(3, it.buf);
    }
    // 4: RemovedIndexes -- ConcurrentStack<int> enumerates from the top of the stack down [inferred: protobuf-net
    // writes a collection in enumeration order]
    if (packed) {
        pb::Writer a; for (auto it = ix.RemovedIndexes.rbegin(); it != ix.RemovedIndexes.rend(); ++it) a.varint((uint64_t)(int64_t)*it);
        if (!ix.RemovedIndexes.empty()) data.bytes_field(4, a.buf);
    } else {
        for (auto it = ix.RemovedIndexes.rbegin(); it != ix.RemovedIndexes.rend(); ++it) data.int32_field(4, *it);
    }
    data.int32_field_nz(5, ix.EntryPointId);
    data.int32_field_nz(6, ix.Capacity);
    data.int32_field_nz(7, ix.Length);
    data.int32_field_nz(8, ix.Count());

    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    pb::Writer head;
    head.tag(1, 2); head.varint(params.buf.size());
    std::fwrite(head.buf.data(), 1, head.buf.size(), f);
    std::fwrite(params.buf.data(), 1, params.buf.size(), f);
    pb::Writer head2;
    head2.tag(2, 2); head2.varint(data.buf.size());
    std::fwrite(head2.buf.data(), 1, head2.buf.size(), f);
    std::fwrite(data.buf.data(), 1, data.buf.size(), f);
    std::fclose(f);
}

// ---- reader -----------------------------------------------------------------
inline EdgeList read_edge_list(pb::Reader r) {
    EdgeList e;
    int count = 0;
    while (!r.done()) {
        uint64_t t = r.varint(); int field = (int)(t >> 3), wt = (int)(t & 7);
        if (field == 1) pb::read_int32s(r, wt, e.Buffer);
        else if (field == 2 && wt == 0) count = (int32_t)r.varint();
        else r.skip(wt);
    }
    if (count > (int)e.Buffer.size()) throw std::runtime_error("snapshot: EdgeList.Count exceeds Buffer");
    e.Count = count;
    return e;
}

inline std::unique_ptr<Index> deserialize(Metric metric, const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> bytes((size_t)sz);
    if (sz > 0 && std::fread(bytes.data(), 1, (size_t)sz, f) != (size_t)sz) { std::fclose(f); throw std::runtime_error("short read"); }
    std::fclose(f);

    Params p;  // field initialisers are the defaults protobuf-net leaves untouched
    bool haveParams = false, haveData = false, anyInEdges = false, haveAllow = false, haveEp = false;
    int ep = -1, capacity = 0, length = 0, count = 0;
    std::vector<std::unique_ptr<Node>> nodes;
    std::vector<int> active, removedTopFirst;
    std::vector<std::vector<float>> items;

    pb::Reader root(bytes.data(), bytes.data() + bytes.size());
    while (!root.done()) {
        uint64_t t = root.varint(); int field = (int)(t >> 3), wt = (int)(t & 7);
        if (field == 1 && wt == 2) {
            haveParams = true;
            pb::Reader r = root.sub();
            while (!r.done()) {
                uint64_t t2 = r.varint(); int f2 = (int)(t2 >> 3), w2 = (int)(t2 & 7);
                if (f2 == 1 && w2 == 0) p.MaxEdges = (int32_t)r.varint();
                else if (f2 == 2 && w2 == 1) { uint64_t b = r.fixed64(); std::memcpy(&p.DistributionRate, &b, 8); }
                else if (f2 == 3 && w2 == 0) p.MinNN = (int32_t)r.varint();
                else if (f2 == 4 && w2 == 0) p.MaxCandidates = (int32_t)r.varint();
                else if (f2 == 5 && w2 == 0) p.RemoveMaxCandidates = (int32_t)r.varint();
                else if (f2 == 6 && w2 == 0) p.CollectionSize = (int32_t)r.varint();
                else if (f2 == 7 && w2 == 0) p.RandomSeed = (int32_t)r.varint();
                else if (f2 == 8 && w2 == 0) { p.AllowRemovals = r.varint() != 0; haveAllow = true; }
                else r.skip(w2);
            }
        } else if (field == 2 && wt == 2) {
            haveData = true;
            pb::Reader r = root.sub();
            while (!r.done()) {
                uint64_t t2 = r.varint(); int f2 = (int)(t2 >> 3), w2 = (int)(t2 & 7);
                if (f2 == 1 && w2 == 2) {
                    pb::Reader nr = r.sub();
                    auto node = std::make_unique<Node>();
                    while (!nr.done()) {
                        uint64_t t3 = nr.varint(); int f3 = (int)(t3 >> 3), w3 = (int)(t3 & 7);
                        if (f3 == 1 && w3 == 0) node->Id = (int32_t)nr.varint();
                        else if (f3 == 2 && w3 == 0) node->IsRemoved = nr.varint() != 0;
                        else if (f3 == 3 && w3 == 2) node->OutEdges.push_back(read_edge_list(nr.sub()));
                        else if (f3 == 4 && w3 == 2) { node->InEdges.push_back(read_edge_list(nr.sub())); anyInEdges = true; }
                        else nr.skip(w3);
                    }
                    nodes.push_back(std::move(node));
                } else if (f2 == 2) pb::read_int32s(r, w2, active);
                else if (f2 == 3 && w2 == 2) {
                    pb::Reader ir = r.sub();
                    std::vector<float> v;
                    while (!ir.done()) {
                        uint64_t t3 = ir.varint(); int f3 = (int)(t3 >> 3), w3 = (int)(t3 & 7);
                        if (f3 == 1) pb::read_floats(ir, w3, v); else ir.skip(w3);
                    }
                    items.push_back(std::move(v));
                } else if (f2 == 4) pb::read_int32s(r, w2, removedTopFirst);
                else if (f2 == 5 && w2 == 0) { ep = (int32_t)r.varint(); haveEp = true; }
                else if (f2 == 6 && w2 == 0) capacity = (int32_t)r.varint();
                else if (f2 == 7 && w2 == 0) length = (int32_t)r.varint();
                else if (f2 == 8 && w2 == 0) count = (int32_t)r.varint();
                else r.skip(w2);
            }
        } else root.skip(wt);
    }
    if (!haveParams) throw std::runtime_error("Parameters cannot be null during deserialization.");  // HNSWIndex.cs:37-38
    if (!haveData) throw std::runtime_error("Data cannot be null during deserialization.");          // HNSWIndex.cs:40-41
    // Latent reference quirks handled deliberately: an omitted
    // zero EntryPointId, and AllowRemovals=false being omitted on write.
    if (!haveEp && !nodes.empty()) ep = 0;
    if (!haveAllow) p.AllowRemovals = anyInEdges || nodes.empty();
    if (nodes.size() != items.size()) throw std::runtime_error("snapshot: nodes/items length mismatch");
    (void)count;

    auto ix = std::make_unique<Index>(metric, p);
    ix->dim = items.empty() ? -1 : (int)items[0].size();
    for (size_t i = 0; i < nodes.size(); i++) {
        if (nodes[i]->Id != (int)i) throw std::runtime_error("snapshot: node ids must equal their slot");
        if ((int)items[i].size() != ix->dim) throw std::runtime_error("snapshot: ragged items");
        ix->Items.insert(ix->Items.end(), items[i].begin(), items[i].end());
    }
    ix->Nodes = std::move(nodes);
    ix->ActiveIds = std::move(active);
    ix->RemovedIndexes.assign(removedTopFirst.rbegin(), removedTopFirst.rend());
    ix->EntryPointId = ep;
    ix->Capacity = capacity > 0 ? capacity : std::max<int>(p.CollectionSize, (int)ix->Nodes.size());
    ix->Length = length > 0 ? length : (int)ix->Nodes.size();
    return ix;
}

}  // namespace hnsw_oracle
