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


def got_paths(root=None):
    root = root or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "got")
    return os.path.join(root, "GOT_Nodes.csv"), os.path.join(root, "GOT_Rels.csv")


def load_got(root=None):
    return load_neo4j_csv(*got_paths(root))
