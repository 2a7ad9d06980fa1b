// snapshot.hpp -- the Serialize / Deserialize file layout of HNSWIndex.Net.
//
// Replaces HNSWIndex.Serialize / Deserialize (HNSWIndex.cs:210-229) and the
// protobuf-net contracts behind them:
//   HNSWIndexSnapshot   {1: Parameters, 2: DataSnapshot}            HNSWIndexSnapshot.cs:12-16
//   HNSWParameters      {1..8}                                      HNSWParameters.cs:12-55
//   GraphDataSnapshot   {1: Nodes, 2: ActiveNodes, 3: Items, 4: RemovedIndexes,
//                        5: EntryPointId, 6: Capacity, 7: Length, 8: Count}   GraphDataSnapshot.cs:13-35
//   Node                {1: Id, 2: IsRemoved, 3: OutEdges[], 4: InEdges[]}    Node.cs:9-25
//   EdgeList            {1: Buffer[], 2: Count}                               Node.cs:33-36
//   NestedArrayWrapper  {1: Values[]}                                         NestedListWrapper.cs:14-20
// protobuf-net 3.2.52 is a NuGet dependency that is not vendored in the reference;
// the encoding below is the public protobuf wire format with protobuf-net's
// documented conventions (zero-valued members omitted, repeated scalars unpacked
// unless asked).  The reader accepts packed and unpacked forms and 64-bit lengths.
//
// The host-side graph uses the same flat layout the GPU uses (see common.cuh).
#pragma once
#include <stdint.h>

#include <cmath>
#include <string>
#include <vector>

namespace hnsw {

// HNSWParameters.cs:13-55
struct Parameters {
    int MaxEdges = 16;
    double DistributionRate = 1.0 / std::log(16.0);
    int MinNN = 5;
    int MaxCandidates = 100;
    int RemoveMaxCandidates = 100;
    int CollectionSize = 65536;
    int RandomSeed = 31337;
    bool AllowRemovals = true;
};

struct HostGraph {
    Parameters params;
    int dim = 0;
    int length = 0;    // slots in use (GraphData.Length)
    int capacity = 0;  // GraphData.Capacity
    int entry = -1;    // EntryPointId
    int stride0 = 0, strideU = 0;
    std::vector<float> vectors;   // [length][dim] dense
    std::vector<int> levels;      // [length] MaxLayer of each node
    std::vector<int> adj0;        // [length][stride0], -1 padded
    std::vector<int> upper_off;   // [length] first upper row of the node or -1
    std::vector<int> adjU;        // [rowsU][strideU], -1 padded
    // removal state (all empty when nothing was ever removed)
    std::vector<unsigned char> removed;  // [length] Node.IsRemoved
    std::vector<int> active;             // ActiveSet.dense[..count] in its own order; empty = 0..length-1
    std::vector<int> removed_stack;      // GraphData.RemovedIndexes, bottom first (back() is handed out next)
};

inline int stride_for(int max_edges) { return ((max_edges < 1 ? 1 : max_edges) + 3) & ~3; }

// Throws std::runtime_error on I/O or format errors.
void write_snapshot(const HostGraph& g, const std::string& path, bool packed);
void read_snapshot(const std::string& path, HostGraph& g);

}  // namespace hnsw
