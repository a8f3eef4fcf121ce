"""A rooted Newick reader for the placed genome tree (`storage/tree/concatenated.tre`).

The reference reads that file with `dendropy.Tree.get_from_path(..., schema='newick', rooting='force-rooted',
preserve_underscores=True)` (checkm/treeParser.py:482,593) and then only walks it: `parent_node`, `child_nodes()`,
`leaf_nodes()`, `sister_nodes()`, `is_internal()`, `label` (internal nodes; pplacer-inserted nodes have none),
`taxon.label` (leaves) and `find_node_with_taxon_label`.  This module gives those names over a flat, iteratively parsed
tree (the genome tree has thousands of leaves and pplacer output can nest deeply: no recursion), with a label index so a
bin lookup is O(1) instead of a scan of every node per bin.

Label rules as in the Newick standard with underscores kept: unquoted labels end at any of `()[],:;` or white space;
quoted labels are delimited by single quotes with `''` standing for one quote; `[...]` comments are skipped; the number
after `:` is the edge length."""


class Taxon(object):
    __slots__ = ('label',)

    def __init__(self, label):
        self.label = label


class Node(object):
    __slots__ = ('label', 'taxon', 'parent_node', '_children', 'edge_length')

    def __init__(self):
        self.label = None
        self.taxon = None
        self.parent_node = None
        self._children = []
        self.edge_length = None

    def child_nodes(self):
        return list(self._children)

    def is_internal(self):
        return len(self._children) != 0

    def is_leaf(self):
        return len(self._children) == 0

    def sister_nodes(self):
        if self.parent_node is None:
            return []
        return [c for c in self.parent_node._children if c is not self]

    def preorder_iter(self):
        stack = [self]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node._children))

    def leaf_iter(self):
        for node in self.preorder_iter():
            if not node._children:
                yield node

    def leaf_nodes(self):
        return list(self.leaf_iter())


class NewickError(ValueError):
    pass


_BREAK = frozenset('()[],:;\'')


class Tree(object):
    def __init__(self, seed_node):
        self.seed_node = seed_node
        self._by_taxon = {}
        for node in seed_node.preorder_iter():
            if node.taxon is not None and node.taxon.label not in self._by_taxon:
                self._by_taxon[node.taxon.label] = node            # first in preorder, as a scan would find it

    @classmethod
    def get_from_path(cls, path, schema='newick', **_ignored):
        if schema != 'newick':
            raise NewickError('only the newick schema is read')
        with open(path) as handle:
            return cls.get_from_string(handle.read())

    @classmethod
    def get_from_string(cls, text):
        return cls(_parse(text))

    def find_node(self, filter_fn):
        for node in self.seed_node.preorder_iter():
            if filter_fn(node):
                return node
        return None

    def find_node_with_taxon_label(self, label):
        return self._by_taxon.get(label)

    def preorder_node_iter(self):
        return self.seed_node.preorder_iter()

    def leaf_nodes(self):
        return self.seed_node.leaf_nodes()


def _parse(text):
    n = len(text)
    i = 0
    root = Node()
    cur = root                      # the node whose label / length the next token belongs to
    opened = 0
    seen_any = False
    finished = False
    while i < n:
        ch = text[i]
        if ch.isspace():
            i += 1
        elif ch == '[':
            depth = 0
            while i < n:            # comments may nest in practice ([&R] etc. are flat); tolerate both
                if text[i] == '[':
                    depth += 1
                elif text[i] == ']':
                    depth -= 1
                    if depth == 0:
                        break
                i += 1
            if i >= n:
                raise NewickError('unterminated comment')
            i += 1
        elif ch == '(':
            child = Node()
            child.parent_node = cur
            cur._children.append(child)
            cur = child
            opened += 1
            seen_any = True
            i += 1
        elif ch == ',':
            if cur.parent_node is None:
                raise NewickError('"," outside of any clade')
            sib = Node()
            sib.parent_node = cur.parent_node
            cur.parent_node._children.append(sib)
            cur = sib
            i += 1
        elif ch == ')':
            if cur.parent_node is None:
                raise NewickError('unbalanced ")"')
            cur = cur.parent_node
            opened -= 1
            i += 1
        elif ch == ':':
            i += 1
            while i < n and text[i].isspace():
                i += 1
            j = i
            while j < n and text[j] not in _BREAK and not text[j].isspace():
                j += 1
            try:
                cur.edge_length = float(text[i:j])
            except ValueError:
                raise NewickError('bad edge length %r' % text[i:j])
            i = j
        elif ch == ';':
            finished = True
            i += 1
            break                   # one tree per file is what pplacer / guppy write
        else:
            if ch == "'":
                parts = []
                i += 1
                while True:
                    j = text.find("'", i)
                    if j < 0:
                        raise NewickError('unterminated quoted label')
                    parts.append(text[i:j])
                    if j + 1 < n and text[j + 1] == "'":
                        parts.append("'")
                        i = j + 2
                        continue
                    i = j + 1
                    break
                label = ''.join(parts)
            else:
                j = i
                while j < n and text[j] not in _BREAK and not text[j].isspace():
                    j += 1
                label = text[i:j]
                i = j
            seen_any = True
            if cur._children:
                cur.label = label
            else:
                cur.taxon = Taxon(label)
    if opened != 0:
        raise NewickError('unbalanced parentheses')
    if not seen_any or (not finished and not seen_any):
        raise NewickError('no tree found')
    return root
