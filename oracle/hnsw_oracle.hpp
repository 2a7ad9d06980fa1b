// =============================================================================
// oracle/hnsw_oracle.hpp -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
//
// A CPU restatement (C++17, host only) of the HNSWIndex.Net v1.6.0 hot path:
// KnnQuery / BatchKnnQuery, Add, the three built-in fp32 metrics and the
// Serialize/Deserialize layout.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, load or call it;
// the product (hnswindex.net_b200/) never links or executes anything in here.
//
// PARITY STATUS: the reference is managed C# and cannot be built or run in this
// image (no dotnet/mono), and its repository holds no golden vectors.  The
// restatement is pinned against (a) every known-answer window the reference's
// own tests assert for this path (tests/test_oracle_reference_windows.py) and
// (b) published first outputs of .NET's seeded System.Random.  Byte-level
// parity of the protobuf-net file layout is UNPINNED (no .NET-written file
// exists to compare against) -- see DESIGN.md "Oracle".
//
// All file:line citations are relative to /root/reference/src/HNSWIndex/.
// =============================================================================
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORACLE_HAVE_AVX 1
#else
#define ORACLE_HAVE_AVX 0
#endif

namespace hnsw_oracle {

// -----------------------------------------------------------------------------
// .NET BCL System.Random(seed): legacy Knuth subtractive generator
// (Net5CompatSeedImpl / CompatPrng in dotnet/runtime).  Used by
// GraphData.cs:42,216 (level draw) and Tests/Utils.cs:37 (fixtures).
// Known first outputs of Next(): seed 0 -> 1559595546, 1 -> 534011718,
// 42 -> 1434747710 (pinned in tests/test_oracle_dotnet_random.py).
// -----------------------------------------------------------------------------
struct DotNetRandom {
    int seedArray[56];
    int inext = 0, inextp = 21;
    explicit DotNetRandom(int seed) {
        const int IMAX = std::numeric_limits<int>::max();
        int subtraction = (seed == std::numeric_limits<int>::min()) ? IMAX : std::abs(seed);
        int mj = 161803398 - subtraction;
        std::memset(seedArray, 0, sizeof(seedArray));
        seedArray[55] = mj;
        int mk = 1, ii = 0;
        for (int i = 1; i < 55; i++) {
            if ((ii += 21) >= 55) ii -= 55;
            seedArray[ii] = mk;
            mk = mj - mk;
            if (mk < 0) mk += IMAX;
            mj = seedArray[ii];
        }
        for (int k = 1; k < 5; k++) {
            for (int i = 1; i < 56; i++) {
                int n = i + 30;
                if (n >= 55) n -= 55;
                // int32 wrap-around subtraction, as unchecked C#
                seedArray[i] = (int)((uint32_t)seedArray[i] - (uint32_t)seedArray[1 + n]);
                if (seedArray[i] < 0) seedArray[i] += IMAX;
            }
        }
        inext = 0;
        inextp = 21;
    }
    int InternalSample() {
        const int IMAX = std::numeric_limits<int>::max();
        int locINext = inext, locINextp = inextp;
        if (++locINext >= 56) locINext = 1;
        if (++locINextp >= 56) locINextp = 1;
        int retVal = (int)((uint32_t)seedArray[locINext] - (uint32_t)seedArray[locINextp]);
        if (retVal == IMAX) retVal--;
        if (retVal < 0) retVal += IMAX;
        seedArray[locINext] = retVal;
        inext = locINext;
        inextp = locINextp;
        return retVal;
    }
    int Next() { return InternalSample(); }
    double Sample() { return InternalSample() * (1.0 / std::numeric_limits<int>::max()); }
    // .NET 8: redraw while the double rounds to 1.0f.
    float NextSingle() {
        for (;;) {
            float f = (float)Sample();
            if (f < 1.0f) return f;
        }
    }
};

// -----------------------------------------------------------------------------
// Metrics -- exact operation order of the AVX+FMA branches.
//   sq_euclid : Metrics/EuclideanMetric.cs:19-60
//   cosine    : Metrics/CosineMetric.cs:23-49,78-91,145-171
//   ucosine   : Metrics/CosineMetric.cs:103-118,135-141
// Two implementations of each: a scalar 8-lane emulation (portable, the spec)
// and AVX intrinsics (fast, used for the CPU baseline).  tests/ check that the
// two agree bit for bit.  Build with -ffp-contract=off.
// -----------------------------------------------------------------------------
enum Metric { SQ_EUCLID = 0, COSINE = 1, UCOSINE = 2, EUCLID = 3 };

inline float hsum_sq(const float* L) {
    // lower+upper (EuclideanMetric.cs:45-47), then two hadd (:48-49)
    float s0 = L[0] + L[4], s1 = L[1] + L[5], s2 = L[2] + L[6], s3 = L[3] + L[7];
    return (s0 + s1) + (s2 + s3);
}
inline float hsum_cos(const float* L) {
    // HorizontalSum256/128 (CosineMetric.cs:145-171): movehl then shuffle 0x55
    float s0 = L[0] + L[4], s1 = L[1] + L[5], s2 = L[2] + L[6], s3 = L[3] + L[7];
    return (s0 + s2) + (s1 + s3);
}

inline float sq_euclid_scalar(const float* a, const float* b, int len) {
    float L[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n8 = len & ~7;
    for (int i = 0; i < n8; i += 8)
        for (int j = 0; j < 8; ++j) {
            float d = a[i + j] - b[i + j];
            L[j] = std::fmaf(d, d, L[j]);
        }
    float r = hsum_sq(L);
    for (int i = n8; i < len; ++i) {  // EuclideanMetric.cs:53-57 (mul, then add)
        float d = a[i] - b[i];
        float p = d * d;
        r = r + p;
    }
    return r;
}

struct CosAcc { float dot, na, nb; };
inline CosAcc cos_acc_scalar(const float* a, const float* b, int len) {
    float D[8] = {0}, A[8] = {0}, B[8] = {0};
    int i = 0;
    for (; i + 8 <= len; i += 8)
        for (int j = 0; j < 8; ++j) {
            float x = a[i + j], y = b[i + j];
            float p = x * y; D[j] = D[j] + p;
            float pa = x * x; A[j] = A[j] + pa;
            float pb = y * y; B[j] = B[j] + pb;
        }
    CosAcc r{hsum_cos(D), hsum_cos(A), hsum_cos(B)};
    for (; i < len; ++i) {  // CosineMetric.cs:78-85
        float x = a[i], y = b[i];
        float p = x * y; r.dot = r.dot + p;
        float pa = x * x; r.na = r.na + pa;
        float pb = y * y; r.nb = r.nb + pb;
    }
    return r;
}
inline float cos_finish(float dot, float na, float nb) {
    // CosineMetric.cs:88-91: fp64 sqrt and product, cast, fp32 divide
    float denom = (float)(std::sqrt((double)na) * std::sqrt((double)nb));
    if (denom < 1e-30f) return 1.0f;
    return 1.0f - dot / denom;
}
inline float cosine_scalar(const float* a, const float* b, int len) {
    CosAcc c = cos_acc_scalar(a, b, len);
    return cos_finish(c.dot, c.na, c.nb);
}
inline float ucosine_scalar(const float* a, const float* b, int len) {
    float D[8] = {0};
    int i = 0;
    for (; i + 8 <= len; i += 8)
        for (int j = 0; j < 8; ++j) {
            float p = a[i + j] * b[i + j];
            D[j] = D[j] + p;
        }
    float dot = hsum_cos(D);
    for (; i < len; ++i) {
        float p = a[i] * b[i];
        dot = dot + p;
    }
    return 1.0f - dot;
}

#if ORACLE_HAVE_AVX
inline float hsum_sq_avx(__m256 acc) {
    __m128 lower = _mm256_castps256_ps128(acc);
    __m128 upper = _mm256_extractf128_ps(acc, 1);
    __m128 s = _mm_add_ps(lower, upper);
    s = _mm_hadd_ps(s, s);
    s = _mm_hadd_ps(s, s);
    return _mm_cvtss_f32(s);
}
inline float hsum_cos_avx(__m256 acc) {
    __m128 s = _mm_add_ps(_mm256_castps256_ps128(acc), _mm256_extractf128_ps(acc, 1));
    __m128 tmp = _mm_add_ps(s, _mm_movehl_ps(s, s));
    __m128 tmp2 = _mm_add_ps(tmp, _mm_shuffle_ps(tmp, tmp, 0x55));
    return _mm_cvtss_f32(tmp2);
}
inline float sq_euclid_avx(const float* a, const float* b, int len) {
    __m256 acc = _mm256_setzero_ps();
    int n8 = len & ~7, i = 0;
    for (; i < n8; i += 8) {
        __m256 d = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        acc = _mm256_fmadd_ps(d, d, acc);
    }
    float r = hsum_sq_avx(acc);
    for (; i < len; ++i) {
        float d = a[i] - b[i];
        float p = d * d;
        r = r + p;
    }
    return r;
}
inline float cosine_avx(const float* a, const float* b, int len) {
    __m256 D = _mm256_setzero_ps(), A = D, B = D;
    int i = 0;
    for (; i + 8 <= len; i += 8) {
        __m256 va = _mm256_loadu_ps(a + i), vb = _mm256_loadu_ps(b + i);
        D = _mm256_add_ps(D, _mm256_mul_ps(va, vb));
        A = _mm256_add_ps(A, _mm256_mul_ps(va, va));
        B = _mm256_add_ps(B, _mm256_mul_ps(vb, vb));
    }
    float dot = hsum_cos_avx(D), na = hsum_cos_avx(A), nb = hsum_cos_avx(B);
    for (; i < len; ++i) {
        float x = a[i], y = b[i];
        float p = x * y; dot = dot + p;
        float pa = x * x; na = na + pa;
        float pb = y * y; nb = nb + pb;
    }
    return cos_finish(dot, na, nb);
}
inline float ucosine_avx(const float* a, const float* b, int len) {
    __m256 D = _mm256_setzero_ps();
    int i = 0;
    for (; i + 8 <= len; i += 8)
        D = _mm256_add_ps(D, _mm256_mul_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i)));
    float dot = hsum_cos_avx(D);
    for (; i < len; ++i) {
        float p = a[i] * b[i];
        dot = dot + p;
    }
    return 1.0f - dot;
}
#endif

inline float metric_scalar(Metric m, const float* a, const float* b, int len) {
    switch (m) {
        case SQ_EUCLID: return sq_euclid_scalar(a, b, len);
        case EUCLID: return (float)std::sqrt((double)sq_euclid_scalar(a, b, len));  // EuclideanMetric.cs:103-109
        case COSINE: return cosine_scalar(a, b, len);
        default: return ucosine_scalar(a, b, len);
    }
}
inline float metric_fast(Metric m, const float* a, const float* b, int len) {
#if ORACLE_HAVE_AVX
    switch (m) {
        case SQ_EUCLID: return sq_euclid_avx(a, b, len);
        case EUCLID: return (float)std::sqrt((double)sq_euclid_avx(a, b, len));
        case COSINE: return cosine_avx(a, b, len);
        default: return ucosine_avx(a, b, len);
    }
#else
    return metric_scalar(m, a, b, len);
#endif
}

// -----------------------------------------------------------------------------
// NodeDistance.cs:5-14, DistanceComparer.cs:6-26
// -----------------------------------------------------------------------------
struct NodeDistance {
    int Id;
    float Dist;
};
inline int float_compare_to(float x, float y) {  // System.Single.CompareTo
    if (x < y) return -1;
    if (x > y) return 1;
    if (x == y) return 0;
    if (std::isnan(x)) return std::isnan(y) ? 0 : -1;
    return 1;
}
struct FartherFirst {  // DistanceComparer: max-heap on Dist
    static int Compare(const NodeDistance& x, const NodeDistance& y) {
        if (x.Dist < y.Dist) return -1;
        if (x.Dist > y.Dist) return 1;
        return float_compare_to(x.Dist, y.Dist);
    }
};
struct CloserFirst {  // ReverseDistanceComparer: min-heap on Dist
    static int Compare(const NodeDistance& x, const NodeDistance& y) {
        if (x.Dist > y.Dist) return -1;
        if (x.Dist < y.Dist) return 1;
        return float_compare_to(y.Dist, x.Dist);
    }
};

// BinaryHeap.cs:6-114 -- same sift rules (ties: SiftUp stops on <=0, SiftDown
// prefers the left child unless left<right, stops on <=0); grows x2 unbounded.
template <class Cmp>
struct BinaryHeap {
    std::vector<NodeDistance> buffer;
    int count = 0;
    explicit BinaryHeap(int capacity = 0) { buffer.resize(capacity > 0 ? capacity : 0); }
    int Count() const { return count; }
    void Push(const NodeDistance& item) {
        if (count == (int)buffer.size()) buffer.resize(buffer.empty() ? 16 : buffer.size() * 2);
        SiftUp(count++, item);
    }
    const NodeDistance& Peek() const { return buffer[0]; }
    NodeDistance Pop() {
        NodeDistance result = buffer[0];
        NodeDistance lastLeaf = buffer[--count];
        if (count != 0) SiftDown(0, lastLeaf, count);
        return result;
    }
    void SiftDown(int i, NodeDistance item, int cnt) {
        int half = cnt >> 1;
        while (i < half) {
            int left = (i << 1) + 1, right = left + 1;
            int maxChild = (right < cnt && Cmp::Compare(buffer[left], buffer[right]) < 0) ? right : left;
            if (Cmp::Compare(buffer[maxChild], item) <= 0) break;
            buffer[i] = buffer[maxChild];
            i = maxChild;
        }
        buffer[i] = item;
    }
    void SiftUp(int i, NodeDistance item) {
        while (i > 0) {
            int p = (i - 1) >> 1;
            NodeDistance parent = buffer[p];
            if (Cmp::Compare(item, parent) <= 0) break;
            buffer[i] = parent;
            i = p;
        }
        buffer[i] = item;
    }
    std::vector<NodeDistance> ToArray() const {
        return std::vector<NodeDistance>(buffer.begin(), buffer.begin() + count);
    }
};

// -----------------------------------------------------------------------------
// .NET MemoryExtensions.Sort(Span<T>, comparer) == ArraySortHelper<T>
// IntrospectiveSort (dotnet/runtime, unstable).  Heuristic.cs:22 sorts the
// candidates with it; the exact permutation only matters under equal fp32
// distances.  Restated from the published algorithm; unpinned.
// -----------------------------------------------------------------------------
namespace dotnet_sort {
using T = NodeDistance;
inline int cmp(const T& a, const T& b) { return FartherFirst::Compare(a, b); }
inline void SwapIfGreater(T* k, int i, int j) {
    if (cmp(k[i], k[j]) > 0) std::swap(k[i], k[j]);
}
inline void InsertionSort(T* k, int n) {
    for (int i = 0; i < n - 1; i++) {
        T t = k[i + 1];
        int j = i;
        while (j >= 0 && cmp(t, k[j]) < 0) {
            k[j + 1] = k[j];
            j--;
        }
        k[j + 1] = t;
    }
}
inline void DownHeap(T* k, int i, int n) {
    T d = k[i - 1];
    while (i <= n >> 1) {
        int child = 2 * i;
        if (child < n && cmp(k[child - 1], k[child]) < 0) child++;
        if (!(cmp(d, k[child - 1]) < 0)) break;
        k[i - 1] = k[child - 1];
        i = child;
    }
    k[i - 1] = d;
}
inline void HeapSort(T* k, int n) {
    for (int i = n >> 1; i >= 1; i--) DownHeap(k, i, n);
    for (int i = n; i > 1; i--) {
        std::swap(k[0], k[i - 1]);
        DownHeap(k, 1, i - 1);
    }
}
inline int PickPivotAndPartition(T* k, int n) {
    int hi = n - 1, middle = hi >> 1;
    SwapIfGreater(k, 0, middle);
    SwapIfGreater(k, 0, hi);
    SwapIfGreater(k, middle, hi);
    T pivot = k[middle];
    std::swap(k[middle], k[hi - 1]);
    int left = 0, right = hi - 1;
    while (left < right) {
        while (cmp(k[++left], pivot) < 0) {}
        while (cmp(pivot, k[--right]) < 0) {}
        if (left >= right) break;
        std::swap(k[left], k[right]);
    }
    if (left != hi - 1) std::swap(k[left], k[hi - 1]);
    return left;
}
inline void IntroSort(T* k, int n, int depthLimit) {
    int partitionSize = n;
    while (partitionSize > 1) {
        if (partitionSize <= 16) {
            if (partitionSize == 2) { SwapIfGreater(k, 0, 1); return; }
            if (partitionSize == 3) {
                SwapIfGreater(k, 0, 1); SwapIfGreater(k, 0, 2); SwapIfGreater(k, 1, 2);
                return;
            }
            InsertionSort(k, partitionSize);
            return;
        }
        if (depthLimit == 0) { HeapSort(k, partitionSize); return; }
        depthLimit--;
        int p = PickPivotAndPartition(k, partitionSize);
        IntroSort(k + p + 1, partitionSize - (p + 1), depthLimit);
        partitionSize = p;
    }
}
inline void Sort(T* k, int n) {
    if (n > 1) {
        int lg = 31 - __builtin_clz((unsigned)n);
        IntroSort(k, n, 2 * (lg + 1));
    }
}
}  // namespace dotnet_sort

// -----------------------------------------------------------------------------
// HNSWParameters.cs:13-55
// -----------------------------------------------------------------------------
struct Params {
    int MaxEdges = 16;
    double DistributionRate = 1.0 / std::log(16.0);
    int MinNN = 5;
    int MaxCandidates = 100;
    int RemoveMaxCandidates = 100;
    int CollectionSize = 65536;
    int RandomSeed = 31337;
    bool AllowRemovals = true;
};

// Node.cs:31-107 (EdgeList: backing Buffer with capacity + Count)
struct EdgeList {
    std::vector<int> Buffer;
    int Count = 0;
    EdgeList() {}
    explicit EdgeList(int capacity) : Buffer(capacity, 0), Count(0) {}
    void Add(int value) {  // Node.cs:66-76, Grow :96-106
        int next = Count + 1;
        if ((int)Buffer.size() < next) {
            int current = (int)Buffer.size();
            int newCapacity = current < 16 ? 16 : current * 2;
            if (newCapacity < next) newCapacity = next;
            Buffer.resize(newCapacity, 0);
        }
        Buffer[Count] = value;
        Count = next;
    }
    bool Remove(int value) {  // Node.cs:79-93 (swap-remove)
        for (int i = 0; i < Count; i++)
            if (Buffer[i] == value) {
                int last = --Count;
                if (i != last) Buffer[i] = Buffer[last];
                return true;
            }
        return false;
    }
    static EdgeList Copy(const EdgeList& o) {  // Node.cs:44-49
        EdgeList e(o.Count);
        std::copy(o.Buffer.begin(), o.Buffer.begin() + o.Count, e.Buffer.begin());
        e.Count = o.Count;
        return e;
    }
};

// Node.cs:7-28
struct Node {
    int Id = 0;
    bool IsRemoved = false;
    std::vector<EdgeList> OutEdges, InEdges;
    std::recursive_mutex OutEdgesLock;  // Monitor is re-entrant
    std::mutex InEdgesLock;
    int MaxLayer() const { return (int)OutEdges.size() - 1; }
};

// VisitedListPool.cs:10-67 -- short[] version stamps.
struct VisitedList {
    short CurrVersion = 0;
    std::vector<short> Nodes;
    explicit VisitedList(int n) : Nodes(n, 0) {}
    void Add(int id) {
        if ((size_t)id >= Nodes.size()) Nodes.resize(std::max<size_t>((size_t)id + 1, Nodes.size() * 2), 0);
        Nodes[id] = CurrVersion;
    }
    bool Contains(int id) const {
        if ((size_t)id >= Nodes.size()) return false;
        return Nodes[id] == CurrVersion;
    }
    void Next() {
        CurrVersion++;
        if (CurrVersion == 0) {
            std::fill(Nodes.begin(), Nodes.end(), (short)0);
            CurrVersion++;
        }
    }
};

struct Counters {
    uint64_t ndist = 0;       // distance evaluations
    uint64_t nexp = 0;        // expanded nodes (rows read), all layers
    uint64_t row_words = 0;   // sum over expansions of (deg + 1)
};

// -----------------------------------------------------------------------------
// The index: GraphData.cs + GraphNavigator.cs + GraphConnector.cs + Heuristic.cs
// + HNSWIndex.cs restated for TVector=float[], TDistance=float.
// -----------------------------------------------------------------------------
class Index {
public:
    Params p;
    Metric metric;
    int dim = -1;
    bool use_scalar_metric = false;  // true: 8-lane scalar emulation; false: AVX

    std::vector<std::unique_ptr<Node>> Nodes;
    std::vector<float> Items;  // row-major slab (the reference has one float[] per item)
    std::vector<int> ActiveIds;
    int EntryPointId = -1;
    int Capacity;
    int Length = 0;

    Index(Metric m, const Params& params) : p(params), metric(m), rng(params.RandomSeed < 0 ? (int)(std::random_device{}() & 0x7fffffff) : params.RandomSeed) {
        Capacity = p.CollectionSize;  // GraphData.cs:42,46
    }

    int Count() const { return (int)ActiveIds.size(); }
    const float* item(int id) const { return Items.data() + (size_t)id * dim; }

    // GraphData.cs:247-250
    int MaxEdges(int layer) const { return layer == 0 ? p.MaxEdges * 2 : p.MaxEdges; }
    int GetTopLayer() const { return Nodes[EntryPointId]->MaxLayer(); }

    float dist_vec(const float* a, const float* b, Counters* c) const {
        if (c) c->ndist++;
        return use_scalar_metric ? metric_scalar(metric, a, b, dim) : metric_fast(metric, a, b, dim);
    }
    float Distance(int a, const float* q, Counters* c) const { return dist_vec(item(a), q, c); }
    float Distance(int a, int b, Counters* c) const { return dist_vec(item(a), item(b), c); }

    // GraphData.cs:211-219
    int GetRandomLayer() {
        float random;
        {
            std::lock_guard<std::mutex> g(rngLock);
            random = rng.NextSingle();
        }
        double v = -std::log((double)random) * p.DistributionRate;
        if (!(v < 2147483648.0)) return -1;  // r == 0 -> +inf -> negative int on x64
        return (int)v;
    }

    // GraphData.cs:79-118
    int AddItem(const float* v, int d) {
        int topLayer = GetRandomLayer();
        if (topLayer < 0) return -1;
        return AddItemWithLevel(v, d, topLayer);
    }
    // NewNode, GraphData.cs:224-242
    std::unique_ptr<Node> NewNode(int index, int topLayer) const {
        auto node = std::make_unique<Node>();
        node->Id = index;
        node->OutEdges.resize(topLayer + 1);
        if (p.AllowRemovals) node->InEdges.resize(topLayer + 1);
        for (int layer = 0; layer <= topLayer; layer++) {
            node->OutEdges[layer] = EdgeList(MaxEdges(layer) + 1);
            if (p.AllowRemovals) node->InEdges[layer] = EdgeList(MaxEdges(layer) + 1);
        }
        return node;
    }
    int AddItemWithLevel(const float* v, int d, int topLayer) {
        std::lock_guard<std::mutex> g(indexLock);
        if (dim < 0) dim = d;
        if (d != dim) throw std::runtime_error("dimension mismatch");
        // "Search for empty spot first", GraphData.cs:84-91
        if (p.AllowRemovals && !RemovedIndexes.empty()) {
            int vacantId = RemovedIndexes.back();
            RemovedIndexes.pop_back();
            Nodes[vacantId] = NewNode(vacantId, topLayer);
            std::copy(v, v + dim, Items.begin() + (size_t)vacantId * dim);
            ActiveAdd(vacantId);
            return vacantId;
        }
        int slotId = Length++;
        if (Length > Capacity) Capacity *= 2;
        if (threaded && (Nodes.size() == Nodes.capacity() || Items.size() + dim > Items.capacity()))
            throw std::runtime_error("threaded add needs reserve()");
        Nodes.push_back(NewNode(slotId, topLayer));
        Items.insert(Items.end(), v, v + dim);
        ActiveAdd(slotId);
        return slotId;
    }

    // ActiveSet.cs:69-93 (dense = ActiveIds, sparse built on first use so that deserialised indexes work too)
    void ActiveAdd(int id) {
        if (sparse_valid) {
            if ((int)ActiveSparse.size() <= id) ActiveSparse.resize((size_t)id + 1, 0);
            ActiveSparse[id] = (int)ActiveIds.size();
        }
        ActiveIds.push_back(id);
    }
    void ActiveRemove(int id) {
        if (!sparse_valid) {
            ActiveSparse.assign(Nodes.size(), 0);
            for (size_t i = 0; i < ActiveIds.size(); i++) ActiveSparse[ActiveIds[i]] = (int)i;
            sparse_valid = true;
        }
        int idx = ActiveSparse[id];
        int lastId = ActiveIds.back();
        ActiveIds.pop_back();
        if (idx < (int)ActiveIds.size()) {  // move last into the removed slot
            ActiveIds[idx] = lastId;
            ActiveSparse[lastId] = idx;
        }
    }

    // ------------------------------------------------------------------------------------------------
    // Removal, single-threaded restatement (HNSWIndex.cs:83-100, GraphConnector.cs:53-167,
    // GraphData.cs:124-199).  The neighbourhood locks of GraphLocker.cs and the TryEnterReadLock(0) at
    // GraphConnector.cs:152 only matter under concurrency and are omitted.
    // ------------------------------------------------------------------------------------------------
    std::vector<int> RemovedIndexes;  // ConcurrentStack<int>: back() is the top

    void Remove(int itemIndex, Counters* c = nullptr) {
        if (!p.AllowRemovals) throw std::runtime_error("Removals are disabled in this index instance.");
        if (itemIndex < 0 || itemIndex >= (int)Nodes.size()) throw std::out_of_range("item index");
        RemoveNodeConnections(Nodes[itemIndex].get(), c);
    }

    // GraphConnector.cs:53-67
    void RemoveNodeConnections(Node* item, Counters* c) {
        item->IsRemoved = true;
        for (int layer = item->MaxLayer(); layer >= 0; layer--) {
            ReplaceEntryPointIfNeeded(item, layer);
            RemoveConnectionsAtLayer(item, layer, c);
            if (layer == 0) {  // GraphData.RemoveItem, GraphData.cs:124-128
                RemovedIndexes.push_back(item->Id);
                ActiveRemove(item->Id);
            }
        }
    }

    // GraphConnector.cs:72-85
    void ReplaceEntryPointIfNeeded(Node* removedNode, int layer) {
        if (removedNode->Id != EntryPointId) return;
        if (TryReplaceEntryPoint(layer)) return;
        if (layer > 0) return;
        if (Count() == 1) {
            EntryPointId = -1;
            return;
        }
        ForceReplaceEntryPoint();
    }

    // GraphData.cs:146-166
    bool TryReplaceEntryPoint(int layer) {
        const EdgeList& e = Nodes[EntryPointId]->OutEdges[layer];
        if (e.Count > 0) {
            int replacementId = -1, maxConnections = -1;
            for (int i = 0; i < e.Count; i++) {
                int neighborId = e.Buffer[i];
                int cnt = Nodes[neighborId]->OutEdges[layer].Count;
                if (cnt > maxConnections) {
                    maxConnections = cnt;
                    replacementId = neighborId;
                }
            }
            EntryPointId = replacementId;
            return true;
        }
        return false;
    }

    // GraphData.cs:172-189 (the node being removed is still active here and may be chosen -- as in the reference)
    void ForceReplaceEntryPoint() {
        if (Count() == 0) return;
        int bestLayer = -1, bestId = -1;
        for (int id : ActiveIds) {
            int ml = Nodes[id]->MaxLayer();
            if (ml > bestLayer) {
                bestLayer = ml;
                bestId = id;
            }
        }
        EntryPointId = bestId;
    }

    // GraphConnector.cs:90-167
    void RemoveConnectionsAtLayer(Node* removedNode, int layer, Counters* c) {
        int maxEdges = MaxEdges(layer);
        // DetachOutgoingReferences, GraphConnector.cs:277-288
        {
            const EdgeList& e = removedNode->OutEdges[layer];
            for (int i = 0; i < e.Count; i++) Nodes[e.Buffer[i]]->InEdges[layer].Remove(removedNode->Id);
        }
        std::vector<int> affectedNodes(removedNode->InEdges[layer].Buffer.begin(),
                                       removedNode->InEdges[layer].Buffer.begin() + removedNode->InEdges[layer].Count);
        auto searchCandidates = SearchLayer<true>(removedNode->Id, layer, p.RemoveMaxCandidates, item(removedNode->Id), c, removedNode->Id);
        std::vector<NodeDistance> candidateDistances;
        for (int affectedNodeId : affectedNodes) {
            Node* affectedNode = Nodes[affectedNodeId].get();
            affectedNode->OutEdges[layer].Remove(removedNode->Id);  // RemoveOutEdge, :268-272
            const EdgeList& out = affectedNode->OutEdges[layer];
            std::vector<int> oldIds(out.Buffer.begin(), out.Buffer.begin() + out.Count);
            std::set<int> oldSeen, newSeen;
            candidateDistances.clear();
            for (int id : oldIds) {
                candidateDistances.push_back(NodeDistance{id, Distance(id, affectedNodeId, c)});
                oldSeen.insert(id);
            }
            for (const NodeDistance& sc : searchCandidates) {
                if (sc.Id == affectedNodeId) continue;
                if (oldSeen.count(sc.Id)) continue;
                candidateDistances.push_back(NodeDistance{sc.Id, Distance(sc.Id, affectedNodeId, c)});
            }
            EdgeList newOut = RelativeNeighborPruning(candidateDistances, maxEdges, c);
            for (int j = 0; j < newOut.Count; j++) newSeen.insert(newOut.Buffer[j]);
            for (int oldNeighborId : oldIds) {  // :135-143
                if (newSeen.count(oldNeighborId)) continue;
                affectedNode->OutEdges[layer].Remove(oldNeighborId);
                Nodes[oldNeighborId]->InEdges[layer].Remove(affectedNodeId);
            }
            for (int j = 0; j < newOut.Count; j++) {  // :146-164
                int newNeighborId = newOut.Buffer[j];
                if (oldSeen.count(newNeighborId)) continue;
                Node* newNeighbor = Nodes[newNeighborId].get();
                if (newNeighbor->IsRemoved) continue;
                affectedNode->OutEdges[layer].Add(newNeighborId);
                newNeighbor->InEdges[layer].Add(affectedNodeId);
            }
        }
    }

    // HNSWIndex.cs:55-65
    int Add(const float* v, int d, Counters* c = nullptr) {
        int id = AddItem(v, d);
        if (id == -1) return id;
        std::lock_guard<std::recursive_mutex> g(Nodes[id]->OutEdgesLock);
        ConnectNewNode(id, c);
        return id;
    }

    // HNSWIndex.cs:70-78 with `threads` workers standing in for Parallel.For.
    void AddBatch(const float* v, int count, int d, int* outIds, int threads, Counters* total = nullptr) {
        if (threads <= 1) {
            Counters c;
            for (int i = 0; i < count; i++) outIds[i] = Add(v + (size_t)i * d, d, total ? &c : nullptr);
            if (total) *total = c;
            return;
        }
        if (dim < 0) dim = d;
        Items.reserve(Items.size() + (size_t)count * d);
        Nodes.reserve(Nodes.size() + count);
        ActiveIds.reserve(ActiveIds.size() + count);
        threaded = true;
        std::atomic<int> next{0};
        std::vector<Counters> cs(threads);
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t]() {
                for (;;) {
                    int i = next.fetch_add(1);
                    if (i >= count) break;
                    outIds[i] = Add(v + (size_t)i * d, d, total ? &cs[t] : nullptr);
                }
            });
        for (auto& th : pool) th.join();
        threaded = false;
        if (total) {
            Counters c;
            for (auto& x : cs) { c.ndist += x.ndist; c.nexp += x.nexp; c.row_words += x.row_words; }
            *total = c;
        }
    }

    // GraphConnector.cs:24-47
    void ConnectNewNode(int nodeId, Counters* c) {
        std::unique_lock<std::mutex> ep(entryPointLock);
        if (EntryPointId < 0) {
            EntryPointId = nodeId;
            return;
        }
        Node* currNode = Nodes[nodeId].get();
        if (currNode->MaxLayer() > GetTopLayer()) {
            AddNewConnections(currNode, c);
            EntryPointId = nodeId;
        } else {
            ep.unlock();
            AddNewConnections(currNode, c);
        }
    }

    // GraphConnector.cs:172-181
    void AddNewConnections(Node* currNode, Counters* c) {
        const float* q = item(currNode->Id);
        Node* bestPeer = FindEntryPoint(currNode->MaxLayer(), q, c);
        for (int layer = std::min(currNode->MaxLayer(), GetTopLayer()); layer >= 0; --layer) {
            int next = ConnectAtLayer(currNode, bestPeer, layer, c);
            bestPeer = Nodes[next].get();
        }
    }

    // GraphConnector.cs:187-217
    int ConnectAtLayer(Node* currNode, Node* bestPeer, int layer, Counters* c) {
        auto topCandidates = SearchLayer<true>(bestPeer->Id, layer, p.MaxCandidates, item(currNode->Id), c);
        EdgeList best = RelativeNeighborPruning(topCandidates, MaxEdges(layer), c);
        currNode->OutEdges[layer] = best;
        if (p.AllowRemovals) currNode->InEdges[layer] = EdgeList::Copy(best);
        for (int i = 0; i < best.Count; ++i) {
            int newNeighborId = best.Buffer[i];
            Node* neighbor = Nodes[newNeighborId].get();
            std::lock_guard<std::recursive_mutex> g(neighbor->OutEdgesLock);
            if (p.AllowRemovals) {
                std::lock_guard<std::mutex> gi(neighbor->InEdgesLock);
                neighbor->InEdges[layer].Add(currNode->Id);
            }
            neighbor->OutEdges[layer].Add(currNode->Id);
            if (neighbor->OutEdges[layer].Count > MaxEdges(layer)) PruneOverflow(neighbor, layer, c);
        }
        return best.Buffer[0];  // GraphConnector.cs:216 (heap root when the heuristic early-returned)
    }

    // GraphConnector.cs:222-262
    void PruneOverflow(Node* node, int layer, Counters* c) {
        EdgeList& old = node->OutEdges[layer];
        std::vector<int> oldIds(old.Buffer.begin(), old.Buffer.begin() + old.Count);
        std::vector<NodeDistance> cd(oldIds.size());
        for (size_t i = 0; i < oldIds.size(); i++) cd[i] = NodeDistance{oldIds[i], Distance(oldIds[i], node->Id, c)};
        EdgeList newOut = RelativeNeighborPruning(cd, MaxEdges(layer), c);
        node->OutEdges[layer] = newOut;
        if (!p.AllowRemovals) return;
        for (int id : oldIds) {
            bool keep = false;
            for (int j = 0; j < newOut.Count; j++)
                if (newOut.Buffer[j] == id) { keep = true; break; }
            if (!keep) {
                Node* nb = Nodes[id].get();
                std::lock_guard<std::mutex> gi(nb->InEdgesLock);
                nb->InEdges[layer].Remove(node->Id);
            }
        }
    }

    // Heuristic.cs:11-46
    EdgeList RelativeNeighborPruning(std::vector<NodeDistance>& candidates, int maxEdges, Counters* c) {
        int n = (int)candidates.size();
        if (n < maxEdges) {
            EdgeList ids(n);
            for (int i = 0; i < n; i++) ids.Add(candidates[i].Id);
            return ids;
        }
        std::vector<NodeDistance> result;
        result.reserve(maxEdges + 1);
        dotnet_sort::Sort(candidates.data(), n);
        for (int i = 0; i < n && (int)result.size() < maxEdges; i++) {
            NodeDistance cand = candidates[i];
            bool acceptable = true;
            for (size_t j = 0; j < result.size(); j++)
                if (Distance(result[j].Id, cand.Id, c) < cand.Dist) { acceptable = false; break; }
            if (acceptable) result.push_back(cand);
        }
        EdgeList outIds(maxEdges + 1);
        for (auto& r : result) outIds.Add(r.Id);
        return outIds;
    }

    // GraphNavigator.cs:27-33 / :39-45 (both call the locking FindEntryAtLayer)
    Node* FindEntryPoint(int dstLayer, const float* q, Counters* c) {
        Node* bestPeer = Nodes[EntryPointId].get();
        for (int layer = bestPeer->MaxLayer(); layer > dstLayer; layer--) bestPeer = FindEntryAtLayer(layer, bestPeer, q, c);
        return bestPeer;
    }

    // GraphNavigator.cs:51-82
    Node* FindEntryAtLayer(int layer, Node* startNode, const float* q, Counters* c) {
        Node* cand = startNode;
        float currDist = Distance(cand->Id, q, c);
        bool changed = true;
        std::vector<int> row;
        while (changed) {
            changed = false;
            std::lock_guard<std::recursive_mutex> g(cand->OutEdgesLock);
            const EdgeList& e = cand->OutEdges[layer];
            row.assign(e.Buffer.begin(), e.Buffer.begin() + e.Count);  // span captured at loop entry
            if (c) { c->nexp++; c->row_words += row.size() + 1; }
            for (int id : row) {
                float d = Distance(id, q, c);
                if (d < currDist) {
                    currDist = d;
                    cand = Nodes[id].get();
                    changed = true;
                }
            }
        }
        return cand;
    }

    // GraphNavigator.cs:123-189 (LOCK) / :194-256 (lock-free twin)
    // `excluded`: the id rejected by the filter delegate of GraphConnector.cs:96 (id => id != removedNode.Id), -1 = no filter
    template <bool LOCK>
    std::vector<NodeDistance> SearchLayer(int entryPointId, int layer, int k, const float* q, Counters* c, int excluded = -1) {
        BinaryHeap<FartherFirst> topCandidates(k);
        BinaryHeap<CloserFirst> candidates(k * 2);
        NodeDistance entry{entryPointId, Distance(entryPointId, q, c)};
        float farthest = std::numeric_limits<float>::max();
        if (entryPointId != excluded) {  // GraphNavigator.cs:133-137
            topCandidates.Push(entry);
            farthest = entry.Dist;
        }
        candidates.Push(entry);
        VisitedList* visited = GetFreeVisitedList();
        visited->Add(entryPointId);
        std::vector<int> row;
        while (candidates.Count() > 0) {
            NodeDistance closest = candidates.Pop();
            if (closest.Dist > farthest && topCandidates.Count() >= k) break;
            Node* node = Nodes[closest.Id].get();
            std::unique_lock<std::recursive_mutex> g(node->OutEdgesLock, std::defer_lock);
            if (LOCK) g.lock();
            const EdgeList& e = node->OutEdges[layer];
            const int* ids = e.Buffer.data();
            int cnt = e.Count;
            if (c) { c->nexp++; c->row_words += cnt + 1; }
            for (int i = 0; i < cnt; ++i) {
                int nb = ids[i];
                if (visited->Contains(nb)) continue;
                float d = Distance(nb, q, c);
                if (topCandidates.Count() < k || d < farthest) {
                    NodeDistance sel{nb, d};
                    candidates.Push(sel);
                    if (nb != excluded) topCandidates.Push(sel);  // GraphNavigator.cs:170-171
                    if (topCandidates.Count() > k) topCandidates.Pop();
                    if (topCandidates.Count() > 0) farthest = topCandidates.Peek().Dist;
                }
                visited->Add(nb);
            }
        }
        ReleaseVisitedList(visited);
        return topCandidates.ToArray();
    }

    // GraphNavigator.cs:262-325 (SearchLayerRange; lock-free, stateless search)
    std::vector<NodeDistance> SearchLayerRange(int entryPointId, int layer, float range, const float* q, Counters* c) {
        BinaryHeap<FartherFirst> topCandidates(MaxEdges(layer));
        BinaryHeap<CloserFirst> candidates(MaxEdges(layer) * 2);
        NodeDistance entry{entryPointId, Distance(entryPointId, q, c)};
        float farthest = std::numeric_limits<float>::max();
        if (entry.Dist <= range) {
            topCandidates.Push(entry);
            farthest = entry.Dist;
        }
        candidates.Push(entry);
        VisitedList* visited = GetFreeVisitedList();
        visited->Add(entryPointId);
        while (candidates.Count() > 0) {
            NodeDistance closest = candidates.Peek();
            if (closest.Dist > farthest && closest.Dist > range) break;
            candidates.Pop();
            const EdgeList& e = Nodes[closest.Id]->OutEdges[layer];
            if (c) { c->nexp++; c->row_words += e.Count + 1; }
            for (int i = 0; i < e.Count; ++i) {
                int nb = e.Buffer[i];
                if (visited->Contains(nb)) continue;
                float d = Distance(nb, q, c);
                if (d <= range) {
                    NodeDistance sel{nb, d};
                    candidates.Push(sel);
                    topCandidates.Push(sel);
                    if (topCandidates.Peek().Dist > range) topCandidates.Pop();
                    if (topCandidates.Count() > 0) farthest = topCandidates.Peek().Dist;
                }
                visited->Add(nb);
            }
        }
        ReleaseVisitedList(visited);
        return topCandidates.ToArray();
    }

    // HNSWIndex.cs:144-155
    std::vector<NodeDistance> RangeQuery(const float* q, float range, Counters* c = nullptr) {
        if (Count() <= 0) return {};
        Node* ep = FindEntryPoint(0, q, c);
        auto top = SearchLayerRange(ep->Id, 0, range, q, c);
        std::stable_sort(top.begin(), top.end(), [](const NodeDistance& a, const NodeDistance& b) { return float_compare_to(a.Dist, b.Dist) < 0; });
        return top;
    }

    // HNSWIndex.cs:107-124 (layer = the optional 4th argument, 0 = base layer)
    std::vector<NodeDistance> KnnQuery(const float* q, int k, Counters* c = nullptr, int layer = 0) {
        if (Count() <= 0 || k < 1) return {};
        int neighborsAmount = std::max(p.MinNN, k);
        Node* ep = FindEntryPoint(layer, q, c);
        if (layer < 0 || layer > ep->MaxLayer()) throw std::out_of_range("layer");  // OutEdges[layer]: IndexOutOfRangeException
        auto top = SearchLayer<false>(ep->Id, layer, neighborsAmount, q, c);
        // OrderBy(c => c.Dist) is a stable sort; Comparer<float>.Default
        std::stable_sort(top.begin(), top.end(), [](const NodeDistance& a, const NodeDistance& b) { return float_compare_to(a.Dist, b.Dist) < 0; });
        if (k < (int)top.size()) top.resize(k);
        return top;
    }

    // HNSWIndex.cs:173-187: one search of width k per layer from min(ep.MaxLayer, maxLayer) down to minLayer, each
    // entered from the closest result of the layer above; a layer's result is the sorted list WITHOUT its first
    // element.  result[layer] stays empty for layers below minLayer.
    std::vector<std::vector<NodeDistance>> MultiLayerKnnQuery(const float* q, int k, int maxLayer, int minLayer) {
        if (Count() <= 0 || k < 1) return {};
        Node* entry = Nodes[EntryPointId].get();
        Node* ep = entry->MaxLayer() >= maxLayer ? FindEntryPoint(maxLayer, q, nullptr) : entry;
        int topL = std::min(ep->MaxLayer(), maxLayer);
        std::vector<std::vector<NodeDistance>> result((size_t)std::max(topL + 1, 0));
        for (int layer = topL; layer >= minLayer; layer--) {
            auto cand = SearchLayer<false>(ep->Id, layer, k, q, nullptr);
            std::stable_sort(cand.begin(), cand.end(), [](const NodeDistance& a, const NodeDistance& b) { return float_compare_to(a.Dist, b.Dist) < 0; });
            ep = Nodes[cand[0].Id].get();
            result[layer].assign(cand.begin() + 1, cand.end());
        }
        return result;
    }

    // HNSWIndex.cs:129-137 + HNSWIndexExports.cs:135-145 (padding with -1 / NaN)
    void BatchKnnQuery(const float* q, int count, int k, int* outIds, float* outDists, int threads, Counters* total = nullptr, int layer = 0) {
        std::atomic<int> next{0};
        std::vector<Counters> cs(std::max(threads, 1));
        auto work = [&](int t) {
            for (;;) {
                int base = next.fetch_add(16);
                if (base >= count) break;
                for (int i = base; i < std::min(count, base + 16); i++) {
                    auto res = KnnQuery(q + (size_t)i * dim, k, total ? &cs[t] : nullptr, layer);
                    int n = std::min(k, (int)res.size());
                    for (int j = 0; j < n; j++) {
                        outIds[(size_t)i * k + j] = res[j].Id;
                        outDists[(size_t)i * k + j] = res[j].Dist;
                    }
                    for (int j = n; j < k; j++) {
                        outIds[(size_t)i * k + j] = -1;
                        outDists[(size_t)i * k + j] = std::numeric_limits<float>::quiet_NaN();
                    }
                }
            }
        };
        if (threads <= 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
            for (auto& th : pool) th.join();
        }
        if (total) {
            Counters c;
            for (auto& x : cs) { c.ndist += x.ndist; c.nexp += x.nexp; c.row_words += x.row_words; }
            *total = c;
        }
    }

    // Exact brute-force top-k with the same metric arithmetic (recall ground truth).
    void ExactKnn(const float* q, int count, int k, int* outIds, float* outDists, int threads) const {
        std::atomic<int> next{0};
        auto work = [&]() {
            std::vector<NodeDistance> all;
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= count) break;
                all.clear();
                for (int id : ActiveIds) all.push_back(NodeDistance{id, dist_vec(item(id), q + (size_t)i * dim, nullptr)});
                int kk = std::min(k, (int)all.size());
                std::partial_sort(all.begin(), all.begin() + kk, all.end(), [](const NodeDistance& a, const NodeDistance& b) {
                    return a.Dist < b.Dist || (a.Dist == b.Dist && a.Id < b.Id);
                });
                for (int j = 0; j < kk; j++) { outIds[(size_t)i * k + j] = all[j].Id; outDists[(size_t)i * k + j] = all[j].Dist; }
                for (int j = kk; j < k; j++) { outIds[(size_t)i * k + j] = -1; outDists[(size_t)i * k + j] = std::numeric_limits<float>::quiet_NaN(); }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < std::max(1, threads); t++) pool.emplace_back(work);
        for (auto& th : pool) th.join();
    }

    // -------------------------------------------------------------------------
    // Design aid (not a reference behaviour): sequential simulation of the GPU
    // engine's batched insertion -- every node of a batch searches the graph as
    // frozen before the batch, then reverse links are applied one by one.
    // Used by tests to bound the recall cost of batching before GPU time is spent.
    // -------------------------------------------------------------------------
    void AddBatchFrozen(const float* v, int count, int d, int* outIds, int ratio, int bmax) {
        struct Req { int n, layer, v; };
        std::vector<int> levels(count);
        for (int i = 0; i < count; i++) levels[i] = GetRandomLayer();  // same draw order as serial Add
        int i = 0;
        while (i < count) {
            if (levels[i] < 0) { outIds[i++] = -1; continue; }
            if (EntryPointId < 0) {  // first node: entry point, no links (GraphConnector.cs:28-33)
                outIds[i] = AddItemWithLevel(v + (size_t)i * d, d, levels[i]);
                EntryPointId = outIds[i];
                i++;
                continue;
            }
            int top0 = GetTopLayer();
            int b = std::max(1, std::min(bmax, Count() / std::max(1, ratio)));
            // a node that raises the top layer is inserted alone
            int e = i;
            if (levels[i] > top0) e = i + 1;
            else while (e < count && e - i < b && levels[e] <= top0) e++;
            std::vector<Req> reqs;
            for (int j = i; j < e; j++) {
                if (levels[j] < 0) { outIds[j] = -1; continue; }
                int id = AddItemWithLevel(v + (size_t)j * d, d, levels[j]);
                outIds[j] = id;
                Node* node = Nodes[id].get();
                const float* q = item(id);
                Node* bestPeer = FindEntryPoint(node->MaxLayer(), q, nullptr);
                for (int layer = std::min(node->MaxLayer(), top0); layer >= 0; --layer) {
                    auto cands = SearchLayer<false>(bestPeer->Id, layer, p.MaxCandidates, q, nullptr);
                    EdgeList best = RelativeNeighborPruning(cands, MaxEdges(layer), nullptr);
                    node->OutEdges[layer] = best;
                    if (p.AllowRemovals) node->InEdges[layer] = EdgeList::Copy(best);
                    for (int t = 0; t < best.Count; t++) reqs.push_back(Req{best.Buffer[t], layer, id});
                    bestPeer = Nodes[best.Buffer[0]].get();
                }
            }
            for (auto& r : reqs) {
                Node* nb = Nodes[r.n].get();
                if (p.AllowRemovals) nb->InEdges[r.layer].Add(r.v);
                nb->OutEdges[r.layer].Add(r.v);
                if (nb->OutEdges[r.layer].Count > MaxEdges(r.layer)) PruneOverflow(nb, r.layer, nullptr);
            }
            if (levels[i] > top0) EntryPointId = outIds[i];
            i = e;
        }
    }

private:
    std::mutex indexLock, entryPointLock, rngLock, poolLock;
    DotNetRandom rng;
    bool threaded = false;
    std::vector<int> ActiveSparse;  // ActiveSet.sparse (id -> position in ActiveIds), valid when sparse_valid
    bool sparse_valid = false;
    std::vector<std::unique_ptr<VisitedList>> visitedPool;

    // VisitedListPool.cs:98-141
    VisitedList* GetFreeVisitedList() {
        std::unique_ptr<VisitedList> r;
        {
            std::lock_guard<std::mutex> g(poolLock);
            if (!visitedPool.empty()) { r = std::move(visitedPool.back()); visitedPool.pop_back(); }
        }
        if (!r) r = std::make_unique<VisitedList>(Capacity);
        r->Next();
        return r.release();
    }
    void ReleaseVisitedList(VisitedList* v) {
        std::lock_guard<std::mutex> g(poolLock);
        visitedPool.emplace_back(v);
    }
};

}  // namespace hnsw_oracle
