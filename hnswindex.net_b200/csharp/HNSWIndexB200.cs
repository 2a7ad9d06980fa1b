// HNSWIndexB200.cs -- C# host over the B200 engine's C ABI (include/hnsw_b200.h) via P/Invoke.
//
// Mirrors the public surface of HNSWIndex<float[], float> (src/HNSWIndex/HNSWIndex.cs:20-250 of the
// reference): HNSWParameters, Add, Remove, KnnQuery / BatchKnnQuery (incl. the `layer` argument), RangeQuery /
// BatchRangeQuery, MultiLayerKnnQuery, GetInfo, GetConnectedComponentCounts, Ids, Serialize, Deserialize, Count --
// restricted, as BASELINE.json's north_star says, to the built-in SquaredEuclidean / Cosine / UnitCosine metrics
// (no managed filter delegates: they cannot run on the GPU).  The native library is loaded by the name the
// reference's own NativeAOT shim uses (HNSWIndex.Native), so dropping the engine's
// HNSWIndex.Native.so next to the application is all a deployment changes.
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no dotnet SDK); kept as the integration artefact a
// maintainer adds to src/HNSWIndex/.  It uses only blittable types and DisableRuntimeMarshalling-safe
// signatures, like bindings/HNSWIndex.Native/HNSWIndexExports.cs.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using System.Text;

namespace HNSWIndex.B200
{
    public enum Metric { SquaredEuclidean, Cosine, UnitCosine }

    /// <summary>Same fields and defaults as HNSWParameters&lt;float&gt; (HNSWParameters.cs:13-55).</summary>
    public sealed class HNSWParameters
    {
        public int MaxEdges { get; set; } = 16;
        public double DistributionRate { get; set; } = 1 / Math.Log(16);
        public int MinNN { get; set; } = 5;
        public int MaxCandidates { get; set; } = 100;
        public int RemoveMaxCandidates { get; set; } = 100;
        public int CollectionSize { get; set; } = 65536;
        public int RandomSeed { get; set; } = 31337;
        public bool AllowRemovals = true;
    }

    public readonly struct KNNResult
    {
        public readonly int Id;
        public readonly float Distance;
        public KNNResult(int id, float distance) { Id = id; Distance = distance; }
    }

    internal static unsafe class Native
    {
        private const string Lib = "HNSWIndex.Native";
        [DllImport(Lib)] internal static extern int hnsw_get_last_error_utf8(byte* buf, int bufLen);
        [DllImport(Lib)] internal static extern nint hnsw_create(byte* metricUtf8);
        [DllImport(Lib)] internal static extern void hnsw_free(nint h);
        [DllImport(Lib)] internal static extern int hnsw_add(nint h, float* vectors, int count, int dim, int* outIds);
        [DllImport(Lib)] internal static extern int hnsw_knn_query(nint h, float* q, int count, int dim, int k, int* outIds, float* outDists);
        [DllImport(Lib)] internal static extern int hnsw_set_collection_size(int v);
        [DllImport(Lib)] internal static extern int hnsw_set_max_edges(int v);
        [DllImport(Lib)] internal static extern int hnsw_set_max_candidates(int v);
        [DllImport(Lib)] internal static extern int hnsw_set_remove_max_candidates(int v);
        [DllImport(Lib)] internal static extern int hnsw_set_distribution_rate(float v);
        [DllImport(Lib)] internal static extern int hnsw_set_random_seed(int v);
        [DllImport(Lib)] internal static extern int hnsw_set_min_nn(int v);
        [DllImport(Lib)] internal static extern int hnsw_set_allow_removals(byte v);
        [DllImport(Lib)] internal static extern int hnsw_count(nint h);
        [DllImport(Lib)] internal static extern int hnsw_dim(nint h);
        [DllImport(Lib)] internal static extern int hnsw_serialize(nint h, byte* pathUtf8, int packed);
        [DllImport(Lib)] internal static extern nint hnsw_deserialize(byte* metricUtf8, byte* pathUtf8);
        [DllImport(Lib)] internal static extern int hnsw_index_set_min_nn(nint h, int minNN);
        [DllImport(Lib)] internal static extern int hnsw_set_device(int device);
        [DllImport(Lib)] internal static extern int hnsw_remove(nint h, int* ids, int count);
        [DllImport(Lib)] internal static extern int hnsw_ids(nint h, int* outIds, int capacity);
        [DllImport(Lib)] internal static extern int hnsw_knn_query_layer(nint h, float* q, int count, int dim, int k, int layer, int* outIds, float* outDists);
        [DllImport(Lib)] internal static extern int hnsw_multilayer_knn_query(nint h, float* q, int count, int dim, int k, int maxLayer, int minLayer, int* outIds, float* outDists);
        [DllImport(Lib)] internal static extern int hnsw_range_query(nint h, float* q, int count, int dim, float range, void** outIds, void** outDists, int* counts);
        [DllImport(Lib)] internal static extern void hnsw_free_results(void** ids, void** dists, int count);
        [DllImport(Lib)] internal static extern int hnsw_get_info(nint h, LayerInfoNative* outInfo, int maxLayers);
        [DllImport(Lib)] internal static extern int hnsw_connected_components(nint h, int* counts, int maxLayers);

        [StructLayout(LayoutKind.Sequential)]
        internal struct LayerInfoNative
        {
            public int LayerId, NodesCount, MaxOutEdges, MinOutEdges, MaxInEdges, MinInEdges;
            public double AvgOutEdges, AvgInEdges;
            public int OutEdgesMedian, InEdgesMedian;
        }

        internal static string LastError()
        {
            int n = hnsw_get_last_error_utf8(null, 0);
            if (n <= 0) return string.Empty;
            var buf = new byte[n + 1];
            fixed (byte* p = buf) hnsw_get_last_error_utf8(p, buf.Length);
            return Encoding.UTF8.GetString(buf, 0, n);
        }
        internal static byte[] Utf8Z(string s) => Encoding.UTF8.GetBytes(s + "\0");
    }

    /// <summary>HNSWIndex&lt;float[], float&gt; on one B200.</summary>
    public sealed unsafe class HNSWIndexB200 : IDisposable
    {
        private static readonly object CreateLock = new();   // the pending parameters are process-global (HNSWIndexExports.cs:16)
        private nint handle;
        private readonly HNSWParameters parameters;
        private int dim = -1;

        private static string MetricName(Metric m) => m switch
        {
            Metric.SquaredEuclidean => "sq_euclid", Metric.Cosine => "cosine", _ => "ucosine",
        };

        public HNSWIndexB200(Metric metric, HNSWParameters? hnswParameters = null)
        {
            parameters = hnswParameters ?? new HNSWParameters();
            lock (CreateLock)
            {
                Native.hnsw_set_max_edges(parameters.MaxEdges);
                Native.hnsw_set_distribution_rate((float)parameters.DistributionRate);
                Native.hnsw_set_min_nn(parameters.MinNN);
                Native.hnsw_set_max_candidates(parameters.MaxCandidates);
                Native.hnsw_set_remove_max_candidates(parameters.RemoveMaxCandidates);
                Native.hnsw_set_collection_size(parameters.CollectionSize);
                Native.hnsw_set_random_seed(parameters.RandomSeed);
                Native.hnsw_set_allow_removals(parameters.AllowRemovals ? (byte)1 : (byte)0);
                fixed (byte* m = Native.Utf8Z(MetricName(metric))) handle = Native.hnsw_create(m);
            }
            if (handle == 0) throw new InvalidOperationException("hnsw_create failed: " + Native.LastError());
        }

        private HNSWIndexB200(nint h) { handle = h; parameters = new HNSWParameters(); dim = Native.hnsw_dim(h); }

        /// <summary>HNSWIndex.Add(TVector) (HNSWIndex.cs:55-65).</summary>
        public int Add(float[] item) => Add(new List<float[]> { item })[0];

        /// <summary>HNSWIndex.Add(List) (HNSWIndex.cs:70-78): one batched GPU insertion.</summary>
        public int[] Add(List<float[]> items)
        {
            if (items.Count == 0) return Array.Empty<int>();
            if (dim < 0) dim = items[0].Length;
            var flat = new float[(long)items.Count * dim];
            for (int i = 0; i < items.Count; i++)
            {
                if (items[i].Length != dim) throw new ArgumentException("all vectors must have the same dimension");
                items[i].CopyTo(flat, (long)i * dim);
            }
            var ids = new int[items.Count];
            int rc;
            fixed (float* v = flat) fixed (int* o = ids) rc = Native.hnsw_add(handle, v, items.Count, dim, o);
            if (rc < 0) throw new InvalidOperationException(Native.LastError());
            return ids;
        }

        /// <summary>HNSWIndex.KnnQuery (HNSWIndex.cs:107-124), layer 0, no filter.</summary>
        public List<KNNResult> KnnQuery(float[] query, int k) => BatchKnnQuery(new[] { query }, k)[0];

        /// <summary>HNSWIndex.BatchKnnQuery (HNSWIndex.cs:129-137): one kernel launch for the whole batch.</summary>
        public List<KNNResult>[] BatchKnnQuery(IList<float[]> queries, int k)
        {
            var result = new List<KNNResult>[queries.Count];
            if (queries.Count == 0 || k < 1) { for (int i = 0; i < result.Length; i++) result[i] = new(); return result; }
            int d = dim < 0 ? queries[0].Length : dim;
            var flat = new float[(long)queries.Count * d];
            for (int i = 0; i < queries.Count; i++) queries[i].CopyTo(flat, (long)i * d);
            var ids = new int[(long)queries.Count * k];
            var dists = new float[(long)queries.Count * k];
            int rc;
            if (parameters.MinNN > 0) Native.hnsw_index_set_min_nn(handle, parameters.MinNN);   // parameters are live in the C# API (HNSWIndex.cs:115)
            fixed (float* q = flat) fixed (int* oi = ids) fixed (float* od = dists)
                rc = Native.hnsw_knn_query(handle, q, queries.Count, d, k, oi, od);
            if (rc < 0) throw new InvalidOperationException(Native.LastError());
            for (int i = 0; i < queries.Count; i++)
            {
                var row = new List<KNNResult>(k);
                for (int j = 0; j < k && ids[(long)i * k + j] >= 0; j++) row.Add(new KNNResult(ids[(long)i * k + j], dists[(long)i * k + j]));
                result[i] = row;
            }
            return result;
        }

        /// <summary>HNSWIndex.Remove(int) / Remove(List) (HNSWIndex.cs:83-100): one batched GPU re-link.</summary>
        public void Remove(int itemIndex) => Remove(new List<int> { itemIndex });
        public void Remove(List<int> indexes)
        {
            if (indexes.Count == 0) return;
            var arr = indexes.ToArray();
            int rc;
            fixed (int* p = arr) rc = Native.hnsw_remove(handle, p, arr.Length);
            if (rc < 0) throw new InvalidOperationException(Native.LastError());
        }

        /// <summary>HNSWIndex.Ids() (HNSWIndex.cs:242-245).</summary>
        public List<int> Ids()
        {
            int n = Native.hnsw_ids(handle, null, 0);
            var ids = new int[Math.Max(n, 1)];
            fixed (int* p = ids) Native.hnsw_ids(handle, p, n);
            return new List<int>(new ArraySegment<int>(ids, 0, n));
        }

        /// <summary>BatchKnnQuery with the reference's `layer` argument (HNSWIndex.cs:129-137).</summary>
        public List<KNNResult>[] BatchKnnQuery(IList<float[]> queries, int k, int layer)
        {
            var result = new List<KNNResult>[queries.Count];
            if (queries.Count == 0 || k < 1) { for (int i = 0; i < result.Length; i++) result[i] = new(); return result; }
            int d = dim < 0 ? queries[0].Length : dim;
            var flat = new float[(long)queries.Count * d];
            for (int i = 0; i < queries.Count; i++) queries[i].CopyTo(flat, (long)i * d);
            var ids = new int[(long)queries.Count * k];
            var dists = new float[(long)queries.Count * k];
            int rc;
            fixed (float* q = flat) fixed (int* oi = ids) fixed (float* od = dists)
                rc = Native.hnsw_knn_query_layer(handle, q, queries.Count, d, k, layer, oi, od);
            if (rc < 0) throw new InvalidOperationException(Native.LastError());
            for (int i = 0; i < queries.Count; i++)
            {
                var row = new List<KNNResult>(k);
                for (int j = 0; j < k && ids[(long)i * k + j] >= 0; j++) row.Add(new KNNResult(ids[(long)i * k + j], dists[(long)i * k + j]));
                result[i] = row;
            }
            return result;
        }

        /// <summary>HNSWIndex.MultiLayerKnnQuery (HNSWIndex.cs:173-187): result[layer] = that layer's neighbours (first entry dropped, as the reference does).</summary>
        public List<KNNResult>[] MultiLayerKnnQuery(float[] query, int k, int maxLayer = int.MaxValue, int minLayer = 0)
        {
            int layers;
            fixed (float* q = query) layers = Native.hnsw_multilayer_knn_query(handle, q, 0, query.Length, k, maxLayer, minLayer, null, null);
            if (layers < 0) throw new InvalidOperationException(Native.LastError());
            var result = new List<KNNResult>[layers];
            if (layers == 0) return result;
            var ids = new int[layers * k];
            var dists = new float[layers * k];
            int rc;
            fixed (float* q = query) fixed (int* oi = ids) fixed (float* od = dists)
                rc = Native.hnsw_multilayer_knn_query(handle, q, 1, query.Length, k, maxLayer, minLayer, oi, od);
            if (rc < 0) throw new InvalidOperationException(Native.LastError());
            for (int l = Math.Max(minLayer, 0); l < layers; l++)
            {
                result[l] = new List<KNNResult>();
                for (int j = 0; j < k && ids[l * k + j] >= 0; j++) result[l].Add(new KNNResult(ids[l * k + j], dists[l * k + j]));
            }
            return result;
        }

        /// <summary>HNSWIndex.BatchRangeQuery (HNSWIndex.cs:160-168).</summary>
        public List<KNNResult>[] BatchRangeQuery(IList<float[]> queries, float range)
        {
            int n = queries.Count;
            var result = new List<KNNResult>[n];
            if (n == 0) return result;
            int d = dim < 0 ? queries[0].Length : dim;
            var flat = new float[(long)n * d];
            for (int i = 0; i < n; i++) queries[i].CopyTo(flat, (long)i * d);
            var idPtrs = new nint[n]; var distPtrs = new nint[n]; var counts = new int[n];
            fixed (float* q = flat) fixed (nint* pi = idPtrs) fixed (nint* pd = distPtrs) fixed (int* pc = counts)
            {
                if (Native.hnsw_range_query(handle, q, n, d, range, (void**)pi, (void**)pd, pc) < 0) throw new InvalidOperationException(Native.LastError());
                for (int i = 0; i < n; i++)
                {
                    var row = new List<KNNResult>(counts[i]);
                    for (int j = 0; j < counts[i]; j++) row.Add(new KNNResult(((int*)idPtrs[i])[j], ((float*)distPtrs[i])[j]));
                    result[i] = row;
                }
                Native.hnsw_free_results((void**)pi, (void**)pd, n);
            }
            return result;
        }
        public List<KNNResult> RangeQuery(float[] query, float range) => BatchRangeQuery(new[] { query }, range)[0];

        /// <summary>HNSWInfo.LayerInfo (HNSWInfo.cs:19-53).</summary>
        public sealed class LayerInfo
        {
            public int LayerId, NodesCount, MaxOutEdges, MinOutEdges, MaxInEdges, MinInEdges, OutEdgesMedian, InEdgesMedian;
            public double AvgOutEdges, AvgInEdges;
        }

        /// <summary>HNSWIndex.GetInfo (HNSWIndex.cs:192-196).</summary>
        public List<LayerInfo> GetInfo()
        {
            int n = Native.hnsw_get_info(handle, null, 0);
            if (n < 0) throw new InvalidOperationException(Native.LastError());
            var raw = new Native.LayerInfoNative[Math.Max(n, 1)];
            fixed (Native.LayerInfoNative* p = raw) Native.hnsw_get_info(handle, p, n);
            var layers = new List<LayerInfo>(n);
            for (int i = 0; i < n; i++)
                layers.Add(new LayerInfo
                {
                    LayerId = raw[i].LayerId, NodesCount = raw[i].NodesCount, MaxOutEdges = raw[i].MaxOutEdges, MinOutEdges = raw[i].MinOutEdges,
                    MaxInEdges = raw[i].MaxInEdges, MinInEdges = raw[i].MinInEdges, AvgOutEdges = raw[i].AvgOutEdges, AvgInEdges = raw[i].AvgInEdges,
                    OutEdgesMedian = raw[i].OutEdgesMedian, InEdgesMedian = raw[i].InEdgesMedian,
                });
            return layers;
        }

        /// <summary>HNSWIndex.GetConnectedComponentCounts (HNSWIndex.cs:202-205).</summary>
        public int[] GetConnectedComponentCounts()
        {
            int n = Native.hnsw_connected_components(handle, null, 0);
            if (n < 0) throw new InvalidOperationException(Native.LastError());
            var counts = new int[Math.Max(n, 0)];
            if (n > 0) fixed (int* p = counts) Native.hnsw_connected_components(handle, p, n);
            return counts;
        }

        /// <summary>HNSWIndex.Serialize (HNSWIndex.cs:210-217): same protobuf-net snapshot layout.</summary>
        public void Serialize(string filePath)
        {
            int rc;
            fixed (byte* p = Native.Utf8Z(filePath)) rc = Native.hnsw_serialize(handle, p, 0);
            if (rc < 0) throw new InvalidOperationException(Native.LastError());
        }

        /// <summary>HNSWIndex.Deserialize (HNSWIndex.cs:222-229).</summary>
        public static HNSWIndexB200 Deserialize(Metric metric, string filePath)
        {
            nint h;
            fixed (byte* m = Native.Utf8Z(MetricName(metric))) fixed (byte* p = Native.Utf8Z(filePath)) h = Native.hnsw_deserialize(m, p);
            if (h == 0) throw new InvalidOperationException("hnsw_deserialize failed: " + Native.LastError());
            return new HNSWIndexB200(h);
        }

        /// <summary>HNSWIndex.Count (HNSWIndex.cs:250).</summary>
        public int Count => Native.hnsw_count(handle);

        public void Dispose()
        {
            if (handle != 0) { Native.hnsw_free(handle); handle = 0; }
            GC.SuppressFinalize(this);
        }
        ~HNSWIndexB200() { if (handle != 0) Native.hnsw_free(handle); }
    }
}
