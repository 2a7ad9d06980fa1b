"""
An INDEPENDENT reader/writer of the snapshot layout for the tests: the five protobuf-net contracts of the reference,
declared as a protobuf schema and handed to google.protobuf (the stock C++/upb implementation shipped in this image).
Nothing here shares code with csrc/snapshot.cpp or oracle/snapshot_codec.hpp.

    HNSWIndexSnapshot   HNSWIndexSnapshot.cs:12-16      HNSWParameters      HNSWParameters.cs:12-55
    GraphDataSnapshot   GraphDataSnapshot.cs:13-35      Node / EdgeList     Node.cs:9-36
    NestedArrayWrapper  NestedListWrapper.cs:14-20

proto3 semantics are exactly protobuf-net's defaults for these contracts: scalar members equal to zero/false are not
written (no [DefaultValue], no IsRequired), and `[packed = false]` reproduces "repeated scalars are unpacked unless
IsPacked = true" (never set in the reference).  The parser accepts both forms, as the protobuf spec requires.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto


def build_messages(packed: bool = False):
    """Message classes of the snapshot schema; `packed` chooses how repeated scalars are WRITTEN (both forms are read)."""
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "hnswsnap_packed.proto" if packed else "hnswsnap.proto"
    fd.package = "hnswsnap" if not packed else "hnswsnap_p"
    fd.syntax = "proto3"
    pk = fd.package

    def fld(msg, name, number, ftype, repeated=False, type_name=None, scalar_list=False):
        f = msg.field.add()
        f.name, f.number, f.type = name, number, ftype
        f.label = F.LABEL_REPEATED if repeated else F.LABEL_OPTIONAL
        if type_name:
            f.type_name = "." + pk + "." + type_name
        if scalar_list:
            f.options.packed = packed

    m = fd.message_type.add(); m.name = "HNSWParameters"
    fld(m, "MaxEdges", 1, F.TYPE_INT32); fld(m, "DistributionRate", 2, F.TYPE_DOUBLE); fld(m, "MinNN", 3, F.TYPE_INT32)
    fld(m, "MaxCandidates", 4, F.TYPE_INT32); fld(m, "RemoveMaxCandidates", 5, F.TYPE_INT32); fld(m, "CollectionSize", 6, F.TYPE_INT32)
    fld(m, "RandomSeed", 7, F.TYPE_INT32); fld(m, "AllowRemovals", 8, F.TYPE_BOOL)

    m = fd.message_type.add(); m.name = "EdgeList"
    fld(m, "Buffer", 1, F.TYPE_INT32, repeated=True, scalar_list=True); fld(m, "Count", 2, F.TYPE_INT32)

    m = fd.message_type.add(); m.name = "Node"
    fld(m, "Id", 1, F.TYPE_INT32); fld(m, "IsRemoved", 2, F.TYPE_BOOL)
    fld(m, "OutEdges", 3, F.TYPE_MESSAGE, repeated=True, type_name="EdgeList")
    fld(m, "InEdges", 4, F.TYPE_MESSAGE, repeated=True, type_name="EdgeList")

    m = fd.message_type.add(); m.name = "NestedArrayWrapper"
    fld(m, "Values", 1, F.TYPE_FLOAT, repeated=True, scalar_list=True)

    m = fd.message_type.add(); m.name = "GraphDataSnapshot"
    fld(m, "Nodes", 1, F.TYPE_MESSAGE, repeated=True, type_name="Node")
    fld(m, "ActiveNodes", 2, F.TYPE_INT32, repeated=True, scalar_list=True)
    fld(m, "Items", 3, F.TYPE_MESSAGE, repeated=True, type_name="NestedArrayWrapper")
    fld(m, "RemovedIndexes", 4, F.TYPE_INT32, repeated=True, scalar_list=True)
    fld(m, "EntryPointId", 5, F.TYPE_INT32); fld(m, "Capacity", 6, F.TYPE_INT32); fld(m, "Length", 7, F.TYPE_INT32); fld(m, "Count", 8, F.TYPE_INT32)

    m = fd.message_type.add(); m.name = "HNSWIndexSnapshot"
    fld(m, "Parameters", 1, F.TYPE_MESSAGE, type_name="HNSWParameters")
    fld(m, "DataSnapshot", 2, F.TYPE_MESSAGE, type_name="GraphDataSnapshot")

    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {name: message_factory.GetMessageClass(pool.FindMessageTypeByName(pk + "." + name))
            for name in ("HNSWParameters", "EdgeList", "Node", "NestedArrayWrapper", "GraphDataSnapshot", "HNSWIndexSnapshot")}


def scan_fields(buf: bytes):
    """Top-level (field number, wire type, payload-or-value) triples of one message body -- a 20-line wire-format walker used
    to assert things the object model hides (field ORDER, packed vs unpacked, presence of zero-valued members)."""
    out, i, n = [], 0, len(buf)

    def varint(i):
        v, s = 0, 0
        while True:
            b = buf[i]; i += 1
            v |= (b & 0x7F) << s; s += 7
            if not b & 0x80:
                return v, i
    while i < n:
        key, i = varint(i)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, i = varint(i); out.append((num, wt, v))
        elif wt == 1:
            out.append((num, wt, buf[i:i + 8])); i += 8
        elif wt == 5:
            out.append((num, wt, buf[i:i + 4])); i += 4
        elif wt == 2:
            ln, i = varint(i); out.append((num, wt, buf[i:i + ln])); i += ln
        else:
            raise ValueError(f"wire type {wt}")
    return out
