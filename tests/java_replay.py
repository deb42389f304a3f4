"""Second, independent restatement of the reference's deterministic loops -- TEST INFRASTRUCTURE ONLY.

`oracle/ppr_oracle.cpp` (mode A) was written from the Java text once; this file was written from the
same text a second time, in plain Python with the Java collection semantics spelled out (FIFO
`ConcurrentLinkedQueue`, membership `HashSet`, `HashMap.get`/`put`/`replace`, `java.util.HashMap`
bucket iteration order, IEEE doubles == Python floats).  `tests/test_java_replay.py` holds the two
against each other bit for bit.  It does not pin the oracle to the reference (the reference still cannot
run here: no JDK, unseeded ThreadLocalRandom) -- it removes single-transcription mistakes from the
parts of the path that are deterministic.  `J/` = /root/reference/src/main/java/joezie/fora_neo4j/.

Edge order: `adjM.forEachRelationship` is the un-vendored HeavyGraph's adjacency order; both
restatements take it to be the CSR order of the export (relationship id == CSV row).
"""
from collections import deque


class JavaHashMap:
    """java.util.HashMap<Long,Double> as far as iteration order goes: power-of-two table, bucket =
    spread(Long.hashCode(key)) & (cap-1), insertion order inside a bucket, resize at 0.75 load keeps the
    relative order of a split bucket (JDK 8 HashMap.resize lo/hi lists).  Buckets stay far below the
    treeify threshold of 8 for the dense small integer keys used here."""

    def __init__(self, cap=16):
        self.cap = cap
        self.size = 0
        self.tab = [[] for _ in range(cap)]

    @staticmethod
    def _hash(key):
        h = (key ^ (key >> 32)) & 0xFFFFFFFF          # Long.hashCode
        return h ^ (h >> 16)                          # HashMap.hash spread

    @classmethod
    def copy_of(cls, other):
        # HashMap(Map m): tableSizeFor(size/0.75 + 1), then put in m's iteration order
        s = other.size
        t = int(s / 0.75 + 1.0)
        cap = 1
        while cap < t:
            cap <<= 1
        m = cls(max(cap, 1))
        for k, v in other.items():
            m.put(k, v)
        return m

    def get(self, key):
        for k, v in self.tab[self._hash(key) & (self.cap - 1)]:
            if k == key:
                return v
        return None

    def put(self, key, val):
        b = self.tab[self._hash(key) & (self.cap - 1)]
        for i, (k, _) in enumerate(b):
            if k == key:
                b[i] = (key, val)
                return
        b.append((key, val))
        self.size += 1
        if self.size > self.cap * 0.75:
            old = self.tab
            self.cap *= 2
            self.tab = [[] for _ in range(self.cap)]
            for bucket in old:
                for k, v in bucket:
                    self.tab[self._hash(k) & (self.cap - 1)].append((k, v))

    def replace(self, key, val):
        if self.get(key) is not None:
            self.put(key, val)

    def clear(self):
        # HashMap.clear keeps the table length
        self.tab = [[] for _ in range(self.cap)]
        self.size = 0

    def items(self):
        for bucket in self.tab:
            for kv in bucket:
                yield kv


def forward_push(rp, ci, s, alpha, rmax):
    """J/Forward_Push.java:63-142.  Returns (reserve, residue, rsum_field, pushes, edges)."""
    n = len(rp) - 1
    deg = [rp[v + 1] - rp[v] for v in range(n)]
    reserve, residue = {}, {}
    rsum_field = None
    if deg[s] == 0:                                               # :72-76
        reserve[s] = 1.0
        return reserve, residue, 0.0, 0, 0
    rsum_local = 1.0
    in_q = {s}
    q = deque([s])
    residue[s] = 1.0
    pushes = edges = 0
    while q:
        v = q.popleft()
        in_q.discard(v)
        x = residue[v]
        residue[v] = 0.0
        reserve[v] = reserve.get(v, 0.0) + x * alpha              # :94
        rsum_local -= x * alpha                                   # :97
        pushes += 1
        if deg[v] == 0:                                           # :101-115 (rsum field NOT refreshed on this path)
            nr = residue[s] + x * (1.0 - alpha)
            residue[s] = nr
            if deg[s] > 0 and nr / float(deg[s]) >= rmax and s not in in_q:
                q.append(s)
                in_q.add(s)
            continue
        avg = ((1.0 - alpha) * x) / float(deg[v])                 # :117
        for e in range(rp[v], rp[v + 1]):
            u = ci[e]
            nr = residue.get(u, 0.0) + avg
            residue[u] = nr
            edges += 1
            du = deg[u]
            ratio = nr / float(du) if du else float("inf")        # Java double division by 0 -> +Infinity
            if ratio >= rmax and u not in in_q:                   # :132
                q.append(u)
                in_q.add(u)
        rsum_field = rsum_local                                   # :140
    return reserve, residue, (1.0 if rsum_field is None else rsum_field), pushes, edges


def backward_search(rp_in, ri_in, outdeg, t, alpha, rmax):
    """J/Backward_Search.java:38-100 on the in-edge lists (rp_in, ri_in).  Returns (reserve, pushes, edges)."""
    reserve, residue = {}, {}
    if rp_in[t + 1] - rp_in[t] == 0:                              # :46-49
        reserve[t] = 1.0
        return reserve, 0, 0
    in_q = {t}
    q = deque([t])
    residue[t] = 1.0
    pushes = edges = 0
    while q:
        v = q.popleft()
        in_q.discard(v)
        x = residue[v]
        residue[v] = 0.0
        reserve[v] = reserve.get(v, 0.0) + x * alpha              # :66
        pushes += 1
        avg = (1.0 - alpha) * x                                   # :72, divided per in-neighbour below
        for e in range(rp_in[v], rp_in[v + 1]):
            u = ri_in[e]
            nr = residue.get(u, 0.0) + avg / outdeg[u]            # :83
            residue[u] = nr
            edges += 1
            if nr > rmax and u not in in_q:                       # :87 (strict, no degree scaling)
                q.append(u)
                in_q.add(u)
    return reserve, pushes, edges


def power_method(rp, ci, s, alpha, iterations=100):
    """J/Power_Method.java:44-101, including HashMap iteration order of `pairs` (the order in which
    pushes of one round add into `residue`, which decides the last bits)."""
    reserve = {}
    residue = JavaHashMap()
    residue.put(s, 1.0)
    for _ in range(iterations):
        pairs = JavaHashMap.copy_of(residue)                      # :58
        residue.clear()
        for v, x in pairs.items():
            d = rp[v + 1] - rp[v]
            if x > 0:
                reserve[v] = reserve.get(v, 0.0) + x * alpha
                remain = x * (1 - alpha)
                if d == 0:                                        # :74-80
                    old = residue.get(s)
                    residue.put(s, (0.0 if old is None else old) + remain)
                else:
                    avg = remain / d
                    for e in range(rp[v], rp[v + 1]):
                        u = ci[e]
                        old = residue.get(u)
                        residue.put(u, (0.0 if old is None else old) + avg)
    return reserve


def kth_largest(vals, k):
    """J/Algo_Util.java:32-50: the k-th largest value, null when k is outside 0..len.  The random pivot
    does not change the answer; sorted() gives the same one.  k == 0 passes the range check, recurses
    left until the range is empty and then dies in nextInt(0) -- callers never ask for it."""
    if k < 0 or k > len(vals):
        return None
    if k == 0:
        raise ValueError("J/Algo_Util.java:60 nextInt(0): bound must be positive")
    return sorted(vals, reverse=True)[k - 1]


def base_all_pairs(rp, ci, alpha, threshold, k=-1):
    """J/Base_Whole_Graph.java:58-164: every vertex as a target in id order, keep pi(v,t) >= threshold
    under source v in target order; k >= 0 keeps the rows' entries >= the k-th largest (ties included),
    stable-sorted descending.  Returns ({v: [(t, pi)]}, pushes, edges)."""
    n = len(rp) - 1
    outdeg = [rp[v + 1] - rp[v] for v in range(n)]
    cnt = [0] * (n + 1)
    for u in ci:
        cnt[u + 1] += 1
    for v in range(n):
        cnt[v + 1] += cnt[v]
    rp_in, fill, ri_in = cnt, cnt[:-1].copy(), [0] * len(ci)
    for v in range(n):
        for e in range(rp[v], rp[v + 1]):
            ri_in[fill[ci[e]]] = v
            fill[ci[e]] += 1
    rows, pushes, edges = {}, 0, 0
    for t in range(n):
        res, p, e = backward_search(rp_in, ri_in, outdeg, t, alpha, threshold)
        pushes += p
        edges += e
        for v, pi in res.items():
            if pi >= threshold:                                   # :83
                rows.setdefault(v, []).append((t, pi))
    if k >= 0:
        for v, row in rows.items():
            kth = kth_largest([p for _, p in row], k)
            keep = [(t, p) for t, p in row if kth is None or p >= kth]   # :137-141
            rows[v] = sorted(keep, key=lambda tp: -tp[1])         # List.sort is stable
    return rows, pushes, edges


# ------------------------------------------------------------------------------------------------
# The randomised algorithms.  The Java draws from an unseeded ThreadLocalRandom; here the terminal of
# "walk number j from vertex v" comes from an injected function (the oracle's Philox walk in the tests),
# and residue maps are visited in ascending vertex id (the oracle's choice; java.util.HashMap order
# would only permute the additions).  What is replayed is everything around the walks.
import math


def conf_whole(n):
    """J/Algo_Conf.java:44-52."""
    return {"delta": 1.0 / float(n), "pfail": 1.0 / float(n)}


def conf_topk(n, k):
    """J/Algo_Conf.java:71-81 -- `node_amount / k` is an int division."""
    return {"min_delta": 1.0 / float(n), "delta": 1.0 / float(k),
            "pfail": 1.0 / float(n) / float(n) / math.log(n // k)}


def monte_carlo(n, s, eps, walk):
    """J/Monte_Carlo.java:136-158; walk(v, j) -> terminal of the j-th walk from v."""
    c = conf_whole(n)
    omega = 3 * math.log(2 / c["pfail"]) / eps / eps / c["delta"]
    cnt = {}
    i = 1
    while i <= omega:
        t = walk(s, i - 1)
        cnt[t] = cnt.get(t, 0) + 1
        i += 1
    return {v: float(x) / omega for v, x in cnt.items()}, i - 1


def fora(rp, ci, s, alpha, eps, walk, walk_cost, max_rounds=64):
    """J/Fora_Whole_Graph.java:82-146 with the oracle's deterministic stand-in for the wall-clock balance
    test of :93 (modelled push cost = pops + edges so far, modelled walk cost = walk_cost*rsum*omega)."""
    n, m = len(rp) - 1, len(ci)
    c = conf_whole(n)
    rsum_local = 1.0
    rmax_local = eps * math.sqrt(c["delta"] / 3.0 / float(m) / math.log(2.0 / c["pfail"])) / (1.0 - alpha)
    omega_local = (eps + 2.0) * math.log(2.0 / c["pfail"]) / eps / eps / c["delta"]
    cost, rounds, pushes, edges = 0.0, 0, 0, 0
    reserve, residue = {}, {}
    while cost < walk_cost * rsum_local * omega_local and rounds < max_rounds:
        reserve, residue, rsum_field, p, e = forward_push(rp, ci, s, alpha, rmax_local)
        pushes += p
        edges += e
        cost += float(p + e)
        rsum_local = rsum_field * (1 - alpha)                    # :101
        rmax_local /= 2.0
        rounds += 1
    num_random_walk = int(omega_local * rsum_local)              # (long) truncation, value is positive
    walks = 0
    for v in sorted(residue):
        rc = residue[v]
        if not rc > 0:
            continue                                             # omega_i == 0: the Java loop body never runs
        reserve[v] = reserve.get(v, 0.0) + rc * alpha            # :120-125
        rc *= (1.0 - alpha)
        omega_i = int(math.ceil(rc / rsum_local * float(num_random_walk)))
        a_i = rc / rsum_local * float(num_random_walk) / float(omega_i)
        incr = a_i / float(num_random_walk) * rsum_local
        for j in range(omega_i):
            t = walk(v, j)                                       # random_walk_no_zero_hop, :132
            reserve[t] = reserve.get(t, 0.0) + incr
        walks += omega_i
    return reserve, {"pushes": pushes, "edges": edges, "rounds": rounds, "walks": walks,
                     "rmax_final": rmax_local * 2.0, "rsum": rsum_local}


def forward_push_topk(rp, ci, s, alpha, st, q, min_rmax, first, rmax):
    """J/Forward_Push.java:144-250.  st = {"residue": {}, "reserve": {}, "rsum": x, "q_next": deque}."""
    deg = lambda v: rp[v + 1] - rp[v]
    residue, reserve = st["residue"], st["reserve"]
    d_s = deg(s)
    if d_s == 0:
        reserve[s] = 1.0
        st["rsum"] = 0.0
        return 0, 0
    if first:
        residue[s] = 1.0
    st["q_next"] = q_next = deque()
    rsum_local = st["rsum"]
    in_q, in_next = set(q), set()
    pushes = edges = 0

    def ratio(x, d):                                             # Java double division: x/0 -> +Inf, 0/0 -> NaN
        return x / float(d) if d else (float("inf") if x > 0 else float("nan"))

    while q:
        v = q.popleft()
        in_q.discard(v)
        x = residue[v]
        d = deg(v)
        if ratio(x, d) >= rmax:                                  # :172
            residue[v] = 0.0
            reserve[v] = reserve.get(v, 0.0) + x * alpha
            rsum_local -= x * alpha
            pushes += 1
            if d == 0:                                           # :186-207
                ns = residue[s] + x * (1 - alpha)
                residue[s] = ns
                if ns / float(d_s) >= rmax and s not in in_q:
                    q.append(s)
                    in_q.add(s)
                elif ns / float(d_s) >= min_rmax and s not in in_next:
                    in_next.add(s)
                    q_next.append(s)
                continue
            avg = ((1.0 - alpha) * x) / float(d)
            for e in range(rp[v], rp[v + 1]):
                u = ci[e]
                nr = residue.get(u, 0.0) + avg
                residue[u] = nr
                edges += 1
                r = ratio(nr, deg(u))
                if r >= rmax and u not in in_q:
                    q.append(u)
                    in_q.add(u)
                elif r >= min_rmax and u not in in_next:
                    in_next.add(u)
                    q_next.append(u)
        elif ratio(x, d) >= min_rmax and v not in in_next:       # :238
            in_next.add(v)
            q_next.append(v)
    st["rsum"] = rsum_local
    return pushes, edges


def fora_topk(rp, ci, s, alpha, eps, k, walk):
    """J/Fora_Topk.java:102-184 (reading :136 with its missing ';'); walk(v, j, round) -> terminal."""
    n, m = len(rp) - 1, len(ci)
    c = conf_topk(n, k)
    epsilon = eps * 0.5
    delta_local, min_delta, pfail = c["delta"], c["min_delta"], c["pfail"]
    min_rmax = epsilon * math.sqrt(min_delta / 3 / m / math.log(2 / pfail))
    rsum_local = 1.0
    st = {"residue": {}, "reserve": {}, "rsum": rsum_local, "q_next": deque()}
    q = deque([s])
    first, rnd, pushes, edges, walks = True, 0, 0, 0, 0
    reserve = {}
    while delta_local >= min_delta:
        rmax_local = epsilon * math.sqrt(delta_local / 3.0 / float(m) / math.log(2.0 / pfail))
        omega_local = (epsilon + 2.0) * math.log(2.0 / pfail) / epsilon / epsilon / delta_local
        if rp[s + 1] - rp[s] == 0:                               # :128-132
            reserve = {s: 1.0}
            rsum_local = 0.0
            break
        rmax_local *= math.sqrt(float(m) * rmax_local) * 3.0     # :133
        p, e = forward_push_topk(rp, ci, s, alpha, st, q, min_rmax, first, rmax_local)
        pushes += p
        edges += e
        first = False
        rsum_local = st["rsum"]
        reserve = dict(st["reserve"])                            # getReserveCopy
        q = deque(st["q_next"])
        rsum_random_walk = rsum_local * (1.0 - alpha)
        num_random_walk = int(omega_local * rsum_random_walk)
        for v in sorted(st["residue"]):
            rc = st["residue"][v]
            if not rc > 0:
                continue
            omega_i = int(math.ceil(rc * float(num_random_walk)))   # :155, no rsum scaling here
            a_i = rc * float(num_random_walk) / float(omega_i)
            incr = a_i / float(num_random_walk)
            for j in range(omega_i):
                t = walk(v, j, rnd)                              # random_walk (zero hops allowed), :159
                reserve[t] = reserve.get(t, 0.0) + incr
            walks += omega_i
        rnd += 1
        vals = list(reserve.values())
        kth = kth_largest(vals, k)
        kth = 0.0 if kth is None else kth
        if kth >= (1 + epsilon) * delta_local or delta_local <= min_delta:
            break
        delta_local = max(min_delta, delta_local / 4.0)
    return reserve, {"pushes": pushes, "edges": edges, "rounds": rnd, "walks": walks, "rsum": rsum_local}


def random_walk(rp, ci, start, alpha, next_double, next_int, no_zero_hop=False, restart=None):
    """J/Monte_Carlo.java:60-97 (random_walk) and :99-133 (random_walk_no_zero_hop) with the two
    ThreadLocalRandom calls injected: next_double() in [0,1), next_int(bound) in [0,bound).
    Returns (terminal, moves) -- moves counts successful neighbour picks and dangling restarts."""
    restart = start if restart is None else restart
    deg = lambda v: rp[v + 1] - rp[v]
    if deg(start) == 0:                                          # :70-72 / :108-110
        return start, 0
    cur, moves = start, 0
    if no_zero_hop:                                              # :112-114, a mandatory first move
        cur = ci[rp[start] + next_int(deg(start))]
        moves = 1
    while True:
        if next_double() < alpha:                                # :76
            break
        d = deg(cur)
        if d > 0:
            cur = ci[rp[cur] + next_int(d)]                      # :84-85
        else:
            cur = restart                                        # :87-90
        moves += 1
    return cur, moves


def compute_error(err_type, gnd_reserve, gnd_topk_ids, gnd_topk, algo_reserve, algo_topk_ids):
    """J/Gen_Util.java:259-326 on plain dicts / lists: gnd_reserve, algo_reserve = {id: ppr}; gnd_topk = {id: ppr} of the power
    method's top-k; *_topk_ids = id lists sorted by value (possibly longer than k, ties)."""
    if err_type == "PRECISION":                                  # :273-281, divided by the ground-truth list length
        hit = 0.0
        for v in algo_topk_ids:
            if v in gnd_topk_ids:
                hit += 1
        return hit / float(len(gnd_topk_ids))
    if err_type == "NDCG":                                       # :283-301
        zk = dcg = 0.0
        for i in range(1, len(gnd_topk) + 1):
            zk += (math.pow(2.0, gnd_topk[gnd_topk_ids[i - 1]]) - 1.0) / math.log(i + 1.0) / math.log(2.0)
        for i in range(1, len(algo_topk_ids) + 1):
            p = gnd_topk.get(algo_topk_ids[i - 1])
            p = 0.0 if p is None else p
            dcg += (math.pow(2.0, p) - 1.0) / math.log(i + 1.0) / math.log(2.0)
        return dcg / zk
    mx = 0.0                                                     # MAX_ERR :308-321: only ground-truth entries are visited
    for v, real in gnd_reserve.items():
        est = algo_reserve.get(v)
        est = 0.0 if est is None else est
        mx = max(mx, abs(est - real))
    return mx
