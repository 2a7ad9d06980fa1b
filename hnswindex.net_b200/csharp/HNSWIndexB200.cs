// HNSWIndexB200.cs -- C# host over the B200 engine's C ABI (include/hnsw_b200.h) via P/Invoke.
//
// Mirrors the public surface of HNSWIndex<float[], float> (src/HNSWIndex/HNSWIndex.cs:20-250 of the
// reference) for the accelerated path: HNSWParameters, Add, KnnQuery, BatchKnnQuery, Serialize,
// Deserialize, Count -- restricted, as BASELINE.json's north_star says, to the built-in
// SquaredEuclidean / Cosine / UnitCosine metrics.  The native library is loaded by the name the
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
