"""One-time graph export: Neo4j-import CSVs / edge lists -> canonical host CSR.

Replaces the Neo4j -> HeavyGraph load of the reference (`PPR.setupAdjMatrix`,
/root/reference/src/main/java/joezie/fora_neo4j/PPR.java:136-152).  HeavyGraph's neighbour order is
unpinned (the neo4j-graph-algorithms jar is not even declared in pom.xml), so the canonical order
here is: ascending target id inside a row, parallel edges kept, degree = row length.

Node ids: the mapped id of a node is its 0-based row in the nodes CSV, which equals the Neo4j node
id of the imported store (`target/got.db`, verified in SURVEY.md F10).
"""
import csv
import os

import numpy as np


def csr_from_edges(n, src, dst):
    """Canonical CSR (row_ptr int64[n+1], col_idx int32[m]) from an edge list; duplicates kept."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if src.size and (src.min() < 0 or src.max() >= n or dst.min() < 0 or dst.max() >= n):
        raise ValueError("edge endpoint out of range")
    order = np.lexsort((dst, src))
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=row_ptr[1:])
    return row_ptr, dst[order].astype(np.int32)


def load_neo4j_csv(nodes_csv, rels_csv):
    """Read `neo4j-admin import` style CSVs (`:ID,name` / `:START_ID,:END_ID,:TYPE`).

    Returns (row_ptr, col_idx, names).  The UTF-8 BOM of GOT_Nodes.csv is tolerated.
    """
    with open(nodes_csv, newline="", encoding="utf-8-sig") as f:
        rd = csv.reader(f)
        header = next(rd)
        id_col = next(i for i, h in enumerate(header) if h.strip().upper().startswith(":ID") or h.strip().endswith(":ID"))
        name_col = next((i for i, h in enumerate(header) if h.strip() == "name"), id_col)
        ids, names = [], []
        for row in rd:
            if not row:
                continue
            ids.append(row[id_col])
            names.append(row[name_col])
    index = {k: i for i, k in enumerate(ids)}
    if len(index) != len(ids):
        raise ValueError("duplicate node ids in " + nodes_csv)
    src, dst = [], []
    with open(rels_csv, newline="", encoding="utf-8-sig") as f:
        rd = csv.reader(f)
        header = [h.strip() for h in next(rd)]
        s_col, e_col = header.index(":START_ID"), header.index(":END_ID")
        for row in rd:
            if not row:
                continue
            src.append(index[row[s_col]])
            dst.append(index[row[e_col]])
    row_ptr, col_idx = csr_from_edges(len(ids), src, dst)
    return row_ptr, col_idx, names


def load_edge_list(path, undirected=False, delimiter=None, comments="#%"):
    """Plain edge-list text (SNAP `ca-GrQc.txt` / `com-amazon.ungraph.txt` style: `# comment` lines, one `src dst` pair per
    line, whitespace- or `delimiter`-separated; BlogCatalog / Flickr `edges.csv` with delimiter=",") -- the form in which the
    dissertation's datasets are distributed before `neo4j-admin import` (Dissertation p.36-37).

    Node ids are relabelled to 0..n-1 in ascending order of the original id (the order `neo4j-admin import` assigns node ids
    for a nodes file sorted by id); `undirected=True` stores both directions of every pair, which is how the dissertation
    loads its (undirected) datasets.  Returns (row_ptr, col_idx, original_ids)."""
    src, dst = [], []
    with open(path, encoding="utf-8-sig") as f:
        for line in f:
            line = line.strip()
            if not line or line[0] in comments:
                continue
            parts = line.split(delimiter)
            if len(parts) < 2:
                raise ValueError("not an edge: %r in %s" % (line, path))
            src.append(int(parts[0]))
            dst.append(int(parts[1]))
    src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
    original_ids, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
    s, d = inv[:src.size], inv[src.size:]
    if undirected:
        loops = s == d                                              # a self-loop is one relationship, not two
        s, d = np.concatenate([s, d[~loops]]), np.concatenate([d, s[~loops]])
    row_ptr, col_idx = csr_from_edges(original_ids.size, s, d)
    return row_ptr, col_idx, original_ids


def write_neo4j_csv(row_ptr, col_idx, nodes_csv, rels_csv, names=None, rel_type="Relation"):
    """The inverse of load_neo4j_csv: `:ID,name` / `:START_ID,:END_ID,:TYPE` files for `neo4j-admin import`
    (the step the reference's README describes for building `target/got.db`)."""
    n = len(row_ptr) - 1
    with open(nodes_csv, "w", newline="", encoding="utf-8") as f:
        w = csv.writer(f)
        w.writerow([":ID", "name"])
        for v in range(n):
            w.writerow([v, names[v] if names is not None else v])
    with open(rels_csv, "w", newline="", encoding="utf-8") as f:
        w = csv.writer(f)
        w.writerow([":START_ID", ":END_ID", ":TYPE"])
        for v in range(n):
            for e in range(int(row_ptr[v]), int(row_ptr[v + 1])):
                w.writerow([v, int(col_idx[e]), rel_type])


NODE_RECORD, REL_RECORD = 15, 34   # Neo4j 3.x "standard" record format


def load_neo4j_store(db_dir):
    """Read the topology straight from a Neo4j 3.x store directory (row N3 of SURVEY 8f) -- no JVM needed.

    `neostore.nodestore.db`: 15-byte records, bit 0 of byte 0 = in use; the node id is the record index.
    `neostore.relationshipstore.db`: 34-byte records, byte 0 = in-use bit + bits 32-34 of the first node,
    bytes 1-4 / 5-8 = low 32 bits of first / second node (big endian), bytes 9-12 = type int whose bits
    28-30 carry bits 32-34 of the second node.  Relationship id = record index, so edges come out in the
    order `GraphLoader` scans them (verified on target/got.db: equal to the import CSV).
    Returns (row_ptr, col_idx, original_ids): mapped id i <-> Neo4j id original_ids[i] (ascending), the
    IdMap of `Graph.toMappedNodeId` / `toOriginalNodeId`.
    """
    nodes = np.fromfile(os.path.join(db_dir, "neostore.nodestore.db"), dtype=np.uint8)
    nodes = nodes[: (nodes.size // NODE_RECORD) * NODE_RECORD].reshape(-1, NODE_RECORD)
    original_ids = np.flatnonzero(nodes[:, 0] & 1).astype(np.int64)
    rels = np.fromfile(os.path.join(db_dir, "neostore.relationshipstore.db"), dtype=np.uint8)
    rels = rels[: (rels.size // REL_RECORD) * REL_RECORD].reshape(-1, REL_RECORD)
    rels = rels[(rels[:, 0] & 1) == 1]
    be32 = lambda b: (b[:, 0].astype(np.int64) << 24) | (b[:, 1].astype(np.int64) << 16) | (b[:, 2].astype(np.int64) << 8) | b[:, 3].astype(np.int64)
    first = be32(rels[:, 1:5]) | (((rels[:, 0].astype(np.int64) & 0x0E) >> 1) << 32)
    type_int = be32(rels[:, 9:13])
    second = be32(rels[:, 5:9]) | ((type_int & 0x70000000) << 4)
    mapped = np.full(int(original_ids.max()) + 1 if original_ids.size else 0, -1, np.int64)
    mapped[original_ids] = np.arange(original_ids.size)
    if rels.shape[0] and (first.max() >= mapped.size or second.max() >= mapped.size or (mapped[first] < 0).any() or (mapped[second] < 0).any()):
        raise ValueError("relationship store references a node that is not in use")
    row_ptr, col_idx = csr_from_edges(original_ids.size, mapped[first], mapped[second])
    return row_ptr, col_idx, original_ids


def got_paths(root=None):
    root = root or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "got")
    return os.path.join(root, "GOT_Nodes.csv"), os.path.join(root, "GOT_Rels.csv")


def load_got(root=None):
    return load_neo4j_csv(*got_paths(root))
