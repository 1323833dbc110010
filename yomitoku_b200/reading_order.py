"""Reading order of page elements (paragraphs, tables, figures, the words inside a paragraph): host logic behind
DocumentAnalyzer.aggregate.  Behavioural contract = reference src/yomitoku/reading_order.py:10-216 (+ the two interval
predicates of utils/misc.py:94-125): a precedence graph between elements that overlap along the reading axis and have
no third element strictly between them, walked by a depth-first traversal that starts from the element closest to the
page origin and only emits an element once everything that must precede it has been emitted.  The traversal is
restated on integer ids with explicit lists; its tie rules (stable sorts, first-come children, the way a finished
branch pulls its successors off the stack) follow the reference step by step because the output order IS the result
(tests/test_layout_logic.py pins it against the reference's own file on random layouts).
"""


def _x_overlap(a, b):
    """utils/misc.py:112-125 - the integer x-ranges share more than a point."""
    return min(int(a[2]), int(b[2])) - max(int(a[0]), int(b[0])) > 0


def _y_overlap(a, b, threshold=0.5):
    """utils/misc.py:94-109 - the y-ranges overlap by at least `threshold` of the shorter one."""
    ay1, ay2, by1, by2 = int(a[1]), int(a[3]), int(b[1]), int(b[3])
    shared = max(0, min(ay2, by2) - max(ay1, by1))
    return not (shared / min(ay2 - ay1, by2 - by1) < threshold)


def _blocked(boxes, i, j, axis):
    """True when some third box that overlaps box i across the reading axis lies strictly inside the gap between i and
    j along it (reading_order.py:84-125).  axis 1: vertical gap (top2bottom), axis 0: horizontal gap."""
    lo, hi = axis, axis + 2
    n_lo, n_hi = boxes[i][lo], boxes[i][hi]
    o_lo, o_hi = boxes[j][lo], boxes[j][hi]
    overlap = _x_overlap if axis == 1 else _y_overlap
    for k, s in enumerate(boxes):
        if k == i or k == j or not overlap(s, boxes[i]):
            continue
        s_lo, s_hi = s[lo], s[hi]
        if (n_hi < s_lo < o_lo and n_hi < s_hi < o_lo) or (o_hi < s_lo < n_lo and o_hi < s_hi < n_lo):
            return True
    return False


class _Graph:
    def __init__(self, boxes):
        self.boxes = boxes
        n = len(boxes)
        self.children = [[] for _ in range(n)]
        self.parents = [[] for _ in range(n)]
        self.distance = [0] * n

    def link(self, a, b):
        if b not in self.children[a]:
            self.children[a].append(b)
            self.parents[b].append(a)


def _build(boxes, direction):
    g = _Graph(boxes)
    n = len(boxes)
    max_x = max(b[2] for b in boxes)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            if direction == "top2bottom":
                if _x_overlap(boxes[i], boxes[j]) and not _blocked(boxes, i, j, 1):
                    if boxes[i][1] < boxes[j][1]:
                        g.link(i, j)
                    else:
                        g.link(j, i)
            elif _y_overlap(boxes[i], boxes[j]) and not _blocked(boxes, i, j, 0):
                ti, tj = boxes[i][2], boxes[j][2]
                if direction == "right2left":
                    first, second = (j, i) if ti < tj else (i, j)
                else:
                    first, second = (j, i) if tj < ti else (i, j)
                g.link(first, second)
        b = boxes[i]
        if direction == "top2bottom":
            g.distance[i] = b[0] + b[1]
        elif direction == "right2left":
            g.distance[i] = (max_x - b[2]) + b[1]
        else:
            g.distance[i] = b[0] * 1 + b[1] * 5
    key = 0 if direction == "top2bottom" else 1          # children left-to-right resp. top-to-bottom (stable)
    for i in range(n):
        g.children[i] = sorted(g.children[i], key=lambda c: boxes[c][key])
    return g


def _walk(g, direction):
    """reading_order.py:14-81 on ids."""
    n = len(g.boxes)
    if n == 0:
        return []
    pending = sorted(range(n), key=lambda i: g.distance[i])
    seen = [False] * n
    stack = [pending.pop(0)]
    order, parked = [], []
    sib_key = 0 if direction in "top2bottom" else 1
    while not all(seen):
        while stack:
            cur = stack.pop()
            emitted = False
            if not seen[cur]:
                if all(seen[p] for p in g.parents[cur]):
                    seen[cur] = True
                    order.append(cur)
                    emitted = True
                elif cur not in parked:
                    parked.append(cur)
            if emitted:
                # everything parked gets another chance: pushed newest-first, so the oldest parked node is examined first
                for node in reversed(parked):
                    stack.append(node)
                    parked.remove(node)
            kids = g.children[cur]
            if kids:
                stack.append(cur)
                stack.append(kids.pop(0))
                continue
            # a leaf: its successors that already wait on the stack move to the top, ordered against the reading axis
            moved = []
            for node in stack:                      # (mutating while iterating, as the reference does)
                if cur in g.parents[node]:
                    moved.append(node)
                    stack.remove(node)
            moved.sort(key=lambda c: g.boxes[c][sib_key], reverse=True)
            stack.extend(moved)
        for node in pending:
            if node in parked:
                continue
            stack.append(node)
            pending.remove(node)
            break
        else:
            if not all(seen) and parked:
                node = parked.pop(0)
                seen[node] = True
                order.append(node)
    return order


def prediction_reading_order(elements, direction, img=None):
    """Sets `.order` on every element (objects with `.box`) and returns the list; fewer than two elements are left
    untouched (reading_order.py:193-216)."""
    if len(elements) < 2:
        return elements
    if direction not in ("top2bottom", "right2left", "left2right"):
        raise ValueError("Invalid direction: %s" % direction)
    boxes = [list(e.box) for e in elements]
    for rank, idx in enumerate(_walk(_build(boxes, direction), direction)):
        elements[idx].order = rank
    return elements
