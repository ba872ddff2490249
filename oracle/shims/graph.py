"""Stand-in for the third-party `graph-theory` package, module name `graph`
(pinned graph-theory 2022.4.3 in the reference's poetry.lock:1115-1116; absent here).

TEST INFRASTRUCTURE ONLY (used to run the unmodified reference when recording traces).

PARITY NOTE: the real package's source is not available offline. This restates its
published behaviour for exactly the calls the reference makes
(map_generator.py:218-264, 289-305; parser.py:27-32, 254-276; environment.py:1684-1690):
  * adjacency = insertion-ordered dict of insertion-ordered dicts; add_edge auto-adds nodes,
    bidirectional=True adds the reverse edge right after the forward one;
  * edges() enumerates outer then inner insertion order;
  * breadth_first_search: FIFO over successors, first-discovered predecessor wins;
  * shortest_path: Dijkstra with heap entries (cost, push counter, node, path) and strict-<
    relaxation (for unit weights this is FIFO BFS in successor insertion order).
Pinned by the reference's golden trajectory (tests/test_data/reproducibility_data.py,
3x3 map seed 0) and by tests/test_map_generator.py / tests/test_parser.py invariants.
"""
from collections import deque
from heapq import heappop, heappush


class Graph:
    def __init__(self, from_dict=None, from_list=None):
        self._nodes = {}
        self._edges = {}

    def add_node(self, node_id, obj=None):
        self._nodes[node_id] = obj
        self._edges.setdefault(node_id, {})

    def add_edge(self, node1, node2, value=1, bidirectional=False):
        if node1 not in self._nodes:
            self.add_node(node1)
        if node2 not in self._nodes:
            self.add_node(node2)
        self._edges[node1][node2] = value
        if bidirectional:
            self._edges[node2][node1] = value

    def del_edge(self, node1, node2):
        del self._edges[node1][node2]

    def edges(self, from_node=None):
        if from_node is not None:
            return [(from_node, n2, v) for n2, v in self._edges.get(from_node, {}).items()]
        return [(n1, n2, v) for n1, d in self._edges.items() for n2, v in d.items()]

    def nodes(self, from_node=None):
        if from_node is not None:
            return list(self._edges.get(from_node, {}).keys())
        return list(self._nodes.keys())

    def node(self, node_id):
        return self._nodes.get(node_id)

    def breadth_first_search(self, start, end):
        if start not in self._nodes or end not in self._nodes:
            raise ValueError("unknown node")
        prev = {start: None}
        q = deque([start])
        while q:
            n = q.popleft()
            if n == end:
                path = []
                while n is not None:
                    path.append(n)
                    n = prev[n]
                return path[::-1]
            for m in self._edges.get(n, {}):
                if m not in prev:
                    prev[m] = n
                    q.append(m)
        return []

    def is_connected(self, n1, n2):
        return bool(self.breadth_first_search(n1, n2))

    def shortest_path(self, start, end):
        q, visited, mins = [(0, 0, start, ())], set(), {start: 0}
        i = 1
        while q:
            cost, _, v1, path = heappop(q)
            if v1 in visited:
                continue
            visited.add(v1)
            path = (v1, path)
            if v1 == end:
                out = []
                while path:
                    out.append(path[0])
                    path = path[1]
                return cost, out[::-1]
            for v2, dist in self._edges.get(v1, {}).items():
                if v2 in visited:
                    continue
                nxt = cost + dist
                prev = mins.get(v2)
                if prev is None or nxt < prev:
                    mins[v2] = nxt
                    heappush(q, (nxt, i, v2, path))
                    i += 1
        return float("inf"), []
