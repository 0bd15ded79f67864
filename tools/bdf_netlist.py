#!/usr/bin/env python3
"""Netlist extractor for Quartus block-design files (.bdf, "graphic" 1.4 s-expressions).

TEST INFRASTRUCTURE.  The reference's FPGA top level is a schematic (`FPGA/UA3REO.bdf`), there is no HDL
top (UA3REO.qsf:42,221), so which module output feeds which module input - I <- sin or I <- cos, what
drives each `reset` / `clk_enable` / `clk` - is only recorded as geometry: symbols with port points,
connectors (two end points, optional net-name label) and junction dots.  This tool recovers the netlist
mechanically so that the wiring the golden model assumes is read from the reference itself
(tests/test_hdl_netlist.py) instead of from somebody's reading of a drawing.

Rules (the Block Editor's): two connector end points at the same coordinate are joined; a connector end
point that lies on the interior of another connector is joined when a junction dot sits there; a port or
pin whose connection point coincides with a connector end point belongs to that net; connectors carrying
the same label text are the same net even when they do not touch (connection by name); a pin's own name
names its net.

Usage: python tools/bdf_netlist.py [/root/reference/FPGA/UA3REO.bdf]  -> prints instance.port -> net
"""
import re
import sys
from collections import defaultdict


def _tokens(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    for m in re.finditer(r'\(|\)|"(?:[^"\\]|\\.)*"|[^\s()"]+', text):
        yield m.group(0)


def parse_sexpr(text):
    stack = [[]]
    for t in _tokens(text):
        if t == "(":
            stack.append([])
        elif t == ")":
            top = stack.pop()
            stack[-1].append(top)
        elif t.startswith('"'):
            stack[-1].append(("str", t[1:-1]))
        else:
            stack[-1].append(t)
    return stack[0]


def _items(node, head):
    return [c for c in node[1:] if isinstance(c, list) and c and c[0] == head]


def _pt(node):
    return (int(node[1]), int(node[2]))


def _strs(node):
    return [c[1][1] for c in _items(node, "text") if isinstance(c[1], tuple)]


class Netlist:
    """ports: {(instance, port_name): net_id}; names: {net_id: sorted label list}; pins: {pin_name: net_id}"""

    def __init__(self, path):
        with open(path, "r", errors="replace") as f:
            top = parse_sexpr(f.read())
        parent = {}

        def find(a):
            parent.setdefault(a, a)
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        def union(a, b):
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[ra] = rb

        connectors, junctions, labels = [], [], []
        self.instances = {}
        port_pts, pin_pts = [], []
        for node in top:
            if not isinstance(node, list) or not node:
                continue
            kind = node[0]
            if kind == "connector":
                pts = [_pt(c) for c in _items(node, "pt")]
                connectors.append((pts[0], pts[1]))
                union(("p",) + pts[0], ("p",) + pts[1])
                for s in _strs(node):
                    labels.append((s, pts[0]))
            elif kind == "junction":
                junctions.append(_pt(_items(node, "pt")[0]))
            elif kind == "symbol":
                rect = _items(node, "rect")[0]
                ox, oy = int(rect[1]), int(rect[2])
                texts = _strs(node)
                module, inst = texts[0], texts[1]
                params = {}
                for p in _items(node, "parameter"):
                    vals = [c[1] for c in p[1:] if isinstance(c, tuple)]
                    if len(vals) >= 2:
                        params[vals[0]] = vals[1]
                self.instances[inst] = {"module": module, "params": params, "ports": {}}
                for port in _items(node, "port"):
                    px, py = _pt(_items(port, "pt")[0])
                    name = _strs(port)[0]
                    direction = "input" if _items(port, "input") else ("output" if _items(port, "output") else "bidir")
                    self.instances[inst]["ports"][name] = direction
                    port_pts.append((inst, name, (ox + px, oy + py)))
            elif kind == "pin":
                rect = _items(node, "rect")[0]
                ox, oy = int(rect[1]), int(rect[2])
                name = _strs(node)[1]
                px, py = _pt(_items(node, "pt")[0])
                pin_pts.append((name, (ox + px, oy + py)))
        # junction dots join connector interiors
        for j in junctions:
            for a, b in connectors:
                if (a[0] == b[0] == j[0] and min(a[1], b[1]) <= j[1] <= max(a[1], b[1])) or \
                   (a[1] == b[1] == j[1] and min(a[0], b[0]) <= j[0] <= max(a[0], b[0])):
                    union(("p",) + j, ("p",) + a)
        # a connector END point on another connector's interior also joins (the editor draws a dot there)
        ends = {e for c in connectors for e in c}
        for e in ends:
            for a, b in connectors:
                if e in (a, b):
                    continue
                if (a[0] == b[0] == e[0] and min(a[1], b[1]) < e[1] < max(a[1], b[1])) or \
                   (a[1] == b[1] == e[1] and min(a[0], b[0]) < e[0] < max(a[0], b[0])):
                    if e in junctions:
                        union(("p",) + e, ("p",) + a)
        for s, pt in labels:
            union(("n", s), ("p",) + pt)
        for name, pt in pin_pts:
            union(("n", name), ("p",) + pt)
        self.ports, self.pins = {}, {}
        for inst, name, pt in port_pts:
            self.ports[(inst, name)] = find(("p",) + pt)
        for name, pt in pin_pts:
            self.pins[name] = find(("n", name))
        names = defaultdict(set)
        for key in list(parent):
            if key[0] == "n":
                names[find(key)].add(key[1])
        self.names = {k: sorted(v) for k, v in names.items()}
        self._find = find

    def net_of(self, inst, port_prefix):
        """net id of instance port whose name starts with port_prefix (bus suffix ignored)"""
        hits = [(k, v) for k, v in self.ports.items() if k[0] == inst and re.sub(r"\[.*", "", k[1]) == port_prefix]
        if len(hits) != 1:
            raise KeyError((inst, port_prefix, [h[0] for h in hits]))
        return hits[0][1]

    def label(self, net):
        return "|".join(self.names.get(net, [])) or None

    def drivers(self, net):
        return sorted("%s.%s" % k for k, v in self.ports.items()
                      if v == net and self.instances[k[0]]["ports"][k[1]] == "output")

    def describe(self, inst, port_prefix):
        """'LABEL' if the net is named, else 'INSTANCE.port' of the output that drives it, else None"""
        net = self.net_of(inst, port_prefix)
        lab = self.label(net)
        drv = self.drivers(net)
        if drv:
            return re.sub(r"\[.*", "", drv[0]) + ("" if not lab else " (" + lab + ")")
        return lab


def main(argv):
    path = argv[1] if len(argv) > 1 else "/root/reference/FPGA/UA3REO.bdf"
    nl = Netlist(path)
    for inst in sorted(nl.instances):
        info = nl.instances[inst]
        print("%s : %s %s" % (inst, info["module"], info["params"] or ""))
        for port, direction in info["ports"].items():
            if direction != "input":
                continue
            print("    %-24s <- %s" % (port, nl.describe(inst, re.sub(r"\[.*", "", port))))


if __name__ == "__main__":
    main(sys.argv)
