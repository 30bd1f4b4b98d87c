"""Small Newick reader + Robinson-Foulds comparator used by tests and bench.py.

(BioPython / dendropy / ete3 are not installed.)  Not on the product path: the
library writes Newick from C++ (csrc/host_newick.cpp).
"""


def parse(s):
    """Returns (children, length, name) lists indexed by node id; root is node 0."""
    s = s.strip()
    if s.endswith(";"):
        s = s[:-1]
    children, length, name = [[]], [0.0], [""]
    stack = [0]
    cur = 0
    i, n = 0, len(s)
    # the outermost parenthesis is the root itself
    first = True
    while i < n:
        c = s[i]
        if c == "(":
            if first:
                first = False
            else:
                children.append([]); length.append(0.0); name.append("")
                v = len(children) - 1
                children[stack[-1]].append(v)
                stack.append(v)
            cur = None
            i += 1
        elif c == ",":
            cur = None
            i += 1
        elif c == ")":
            cur = stack.pop() if len(stack) > 1 else 0
            i += 1
        elif c == ":":
            j = i + 1
            while j < n and s[j] not in ",()":
                j += 1
            if cur is not None:
                length[cur] = float(s[i + 1:j])
            i = j
        else:
            j = i
            while j < n and s[j] not in ":,()":
                j += 1
            label = s[i:j].strip()
            if cur is None:
                children.append([]); length.append(0.0); name.append(label)
                cur = len(children) - 1
                children[stack[-1]].append(cur)
            else:
                name[cur] = label
            i = j
    return children, length, name


def _leafsets(children, name):
    """Post-order bitmask of leaves under each node (iterative)."""
    leaf_names = sorted(nm for v, nm in enumerate(name) if not children[v])
    idx = {nm: k for k, nm in enumerate(leaf_names)}
    mask = [0] * len(children)
    order = []
    st = [0]
    while st:
        v = st.pop()
        order.append(v)
        st.extend(children[v])
    for v in reversed(order):
        if not children[v]:
            mask[v] = 1 << idx[name[v]]
        else:
            m = 0
            for c in children[v]:
                m |= mask[c]
            mask[v] = m
    return mask, leaf_names


def bipartitions(s, min_len=None):
    """Set of non-trivial splits (canonical side = the one without leaf 0).

    If min_len is given, internal edges with length <= min_len are collapsed."""
    children, length, name = parse(s)
    mask, leaves = _leafsets(children, name)
    full = (1 << len(leaves)) - 1
    out = set()
    for v in range(1, len(children)):
        if not children[v]:
            continue
        if min_len is not None and length[v] <= min_len:
            continue
        m = mask[v]
        if m & 1:
            m = full ^ m
        if m == 0 or (m & (m - 1)) == 0:
            continue
        out.add(m)
    return out, leaves


def rf_distance(a, b, min_len=None):
    sa, la = bipartitions(a, min_len)
    sb, lb = bipartitions(b, min_len)
    if la != lb:
        raise ValueError("leaf sets differ")
    return len(sa ^ sb)


def branch_lengths_by_split(s):
    """{canonical split: length} incl. pendant edges; root-adjacent edges of a rooted
    binary tree are merged (their split is the same)."""
    children, length, name = parse(s)
    mask, leaves = _leafsets(children, name)
    full = (1 << len(leaves)) - 1
    out = {}
    for v in range(1, len(children)):
        m = mask[v]
        if m & 1:
            m = full ^ m
        out[m] = out.get(m, 0.0) + length[v]
    return out


def max_branch_diff(a, b):
    la, lb = branch_lengths_by_split(a), branch_lengths_by_split(b)
    if set(la) != set(lb):
        return float("inf")
    return max(abs(la[k] - lb[k]) for k in la) if la else 0.0
