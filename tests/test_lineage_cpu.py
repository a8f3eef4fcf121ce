"""Lineage marker-set selection on the placed genome tree (checkm/treeParser.py:468-553 and helpers)
against goldens produced by the REFERENCE's own TreeParser (tests/golden/make_lineage_goldens.py), plus the Newick
reader underneath it.  Host logic only: no GPU."""
import json
import os
import shutil

import pytest

from conftest import GOLDEN

LIN = os.path.join(GOLDEN, 'lineage')


class _Counts(object):
    def __init__(self, pair):
        self.pair = tuple(pair)

    def countUniqueHits(self):
        return self.pair


class _Results(object):
    """What getBinMarkerSets reads from a ResultsParser: results[binId].countUniqueHits() (treeParser.py:536)."""
    def __init__(self, unique_multi):
        self.results = {b: _Counts(p) for b, p in unique_multi.items()}


@pytest.fixture(scope='module')
def expected():
    with open(os.path.join(LIN, 'expected.json')) as f:
        return json.load(f)


@pytest.fixture()
def rundir(tmp_path, expected):
    from checkm_b200.defaultValues import DefaultValues
    saved = DefaultValues.CHECKM_DATA_DIR
    DefaultValues.set_data_root(os.path.join(LIN, 'data'))
    out = str(tmp_path / 'out')
    os.makedirs(os.path.join(out, 'storage', 'tree'))
    shutil.copyfile(os.path.join(LIN, 'tree', 'concatenated.tre'), os.path.join(out, 'storage', 'tree', 'concatenated.tre'))
    for b in expected['unique_multi']:
        os.makedirs(os.path.join(out, 'bins', b))
    yield out
    DefaultValues.set_data_root(saved)


def _parse(path):
    from checkm_b200.markerSets import _parse_set_list
    out = {}
    with open(path) as f:
        assert f.readline() == '# [Lineage Marker File]\n'
        for line in f:
            fields = line.rstrip('\n').split('\t')
            out[fields[0]] = [[fields[2 + 4 * i], fields[3 + 4 * i], int(fields[4 + 4 * i]),
                               [sorted(s) for s in _parse_set_list(fields[5 + 4 * i])]] for i in range(int(fields[1]))]
    return out


def test_marker_file_equals_the_references(expected, rundir, tmp_path):
    from checkm_b200.markerSets import MarkerSetParser, BinMarkerSets
    from checkm_b200.treeParser import TreeParser
    tp = TreeParser()
    rp = _Results(expected['unique_multi'])
    for name, v in expected['variants'].items():
        mf = str(tmp_path / (name + '.ms'))
        ng, boot, noref, force, reqtax, minu, maxm = v['options']
        tp.getBinMarkerSets(rundir, mf, ng, boot, noref, force, reqtax, rp, minu, maxm)
        assert _parse(mf) == v['bins'], name
        # and the file is what the main search's marker-file reader accepts (markerSets.py:490-522)
        msp = MarkerSetParser(1)
        assert msp.markerFileType(mf) == BinMarkerSets.TREE_MARKER_SET
        parsed = msp.parseLineageMarkerSetFile(mf)
        assert {b: [str(ms.UID) for ms in bms.markerSets] for b, bms in parsed.items()} == \
               {b: [s[0] for s in sets] for b, sets in v['bins'].items()}
    # every branch of the selection is in the goldens: the single-genome node is skipped, the unnamed node takes its next
    # named ancestor's name, the bin outside the tree gets the root set, the bin above the domain node starts below it
    d = expected['variants']['lineage_wf_defaults']['bins']
    assert [s[0] for s in d['binA']] == ['UID20', 'UID10', 'UID2', 'UID1']
    assert d['binD'][0][:2] == ['UID22', 'p__Proteobacteria']
    assert [s[0] for s in d['binF']] == ['UID1'] and [s[0] for s in d['binB']] == ['UID2', 'UID1']
    assert [s[0] for s in expected['variants']['few_markers_allowed']['bins']['binG']] == ['UID21', 'UID10', 'UID2', 'UID1']


def test_tree_lookups_equal_the_references(expected, rundir):
    from checkm_b200.markerSets import _parse_set_list
    from checkm_b200.treeParser import TreeParser
    tp = TreeParser()
    bins = sorted(expected['unique_multi'])
    assert tp.getBinTaxonomy(rundir, bins) == expected['bin_taxonomy']
    assert tp.getInsertionBranchId(rundir, bins) == expected['insertion_uid']
    meta = tp.readLineageMetadata(rundir, bins)
    for b in bins:
        got = dict(meta[b])
        if got['marker set'] != 'NA':
            got['marker set'] = [sorted(s) for s in _parse_set_list(got['marker set'])]
        assert {k: repr(v) for k, v in got.items()} == {k: repr(v) for k, v in expected['lineage_metadata'][b].items()}, b


def test_broken_inputs_fail_loudly(rundir, expected, tmp_path):
    from checkm_b200.treeParser import TreeParser
    tre = os.path.join(rundir, 'storage', 'tree', 'concatenated.tre')
    rp = _Results({b: (30, 0) for b in expected['unique_multi']})
    with open(tre, 'w') as f:                                  # UIDX has no line in genome_tree.metadata.tsv
        f.write("((IMG_1,IMG_4):0.1,((binA,IMG_2),IMG_3)'UIDX|k__Bacteria|')'UID1|root|';")
    with pytest.raises(KeyError):
        TreeParser().getBinMarkerSets(rundir, str(tmp_path / 'x.ms'), 2, 0, True, False, False, rp, 10, 10)
    with open(tre, 'w') as f:                                  # a bin under the root with no labelled domain node below it
        f.write("((binA,IMG_1):0.1,IMG_2)'UID1|root|';")
    with pytest.raises(SystemExit):
        TreeParser().getBinMarkerSets(rundir, str(tmp_path / 'y.ms'), 2, 0, True, False, False, rp, 10, 10)
    with open(tre, 'w') as f:                                  # no labelled node above the bin at all
        f.write("((binA,IMG_1):0.1,IMG_2);")
    with pytest.raises(SystemExit):
        TreeParser().getBinMarkerSets(rundir, str(tmp_path / 'z.ms'), 2, 0, True, False, False, rp, 10, 10)


def test_newick_reader():
    from checkm_b200.util import newick
    t = newick.Tree.get_from_string("[&R] ((a_b:0.1,'it''s  here':2e-3)'UID7|p__X y|':1,[c] (c , d)0.93:0.5 , e)root;")
    root = t.seed_node
    assert root.label == 'root' and root.parent_node is None and root.taxon is None
    kids = root.child_nodes()
    assert [k.label for k in kids] == ['UID7|p__X y|', '0.93', None]
    assert kids[2].taxon.label == 'e' and kids[2].is_leaf() and not kids[2].is_internal()
    assert [l.taxon.label for l in root.leaf_nodes()] == ['a_b', "it's  here", 'c', 'd', 'e']       # underscores kept
    assert kids[0].edge_length == 1.0 and kids[0].child_nodes()[1].edge_length == 2e-3
    n = t.find_node_with_taxon_label('d')
    assert n.parent_node is kids[1] and [s.taxon.label for s in n.sister_nodes()] == ['c']
    assert t.find_node_with_taxon_label('nope') is None
    assert t.find_node(lambda x: x.parent_node is None) is root
    # a caterpillar 60,000 levels deep (no recursion anywhere)
    depth = 60000
    deep = newick.Tree.get_from_string('(' * depth + 'x0' + ''.join(',x%d)' % (i + 1) for i in range(depth)) + ';')
    assert len(deep.leaf_nodes()) == depth + 1
    leaf = deep.find_node_with_taxon_label('x0')
    steps = 0
    while leaf.parent_node is not None:
        leaf = leaf.parent_node
        steps += 1
    assert steps == depth
    for bad in ('((a,b);', '(a,b));', "('a,b);", '(a,b):x;', '', '(a[b,c);'):
        with pytest.raises(newick.NewickError):
            newick.Tree.get_from_string(bad)


def test_newick_reader_on_random_trees():
    """Seeded random trees written out with every label style (bare, quoted with blanks / quotes / brackets), optional edge
    lengths and comments: the reader must give back the generating structure."""
    import random
    from checkm_b200.util import newick

    def label(rng, leaf):
        kind = rng.random()
        if kind < 0.5:
            return ('IMG_%d' % rng.randrange(10 ** 6)) if leaf else ('UID%d|k__X;p__Y_%d|%d' % (rng.randrange(999), rng.randrange(99), rng.randrange(101)))
        if kind < 0.8:
            return "bin %d (draft)'s [v2]" % rng.randrange(1000)
        return 'a_b-c.%d' % rng.randrange(1000)

    def quote(s):
        bare = all(c not in s for c in " ()[]':;,")
        return s if bare else "'" + s.replace("'", "''") + "'"

    def build(rng, depth):
        if depth == 0 or rng.random() < 0.3:
            return (label(rng, True), [])
        kids = [build(rng, depth - 1) for _ in range(rng.randrange(2, 5))]
        return (label(rng, False) if rng.random() < 0.6 else None, kids)

    def write(node, rng):
        lab, kids = node
        s = ''
        if kids:
            s += '(' + ','.join(write(k, rng) for k in kids) + ')'
        if lab is not None:
            s += quote(lab)
        if rng.random() < 0.7:
            s += ':%g' % rng.uniform(0, 2)
        if rng.random() < 0.1:
            s += '[&support=%d]' % rng.randrange(100)
        return s

    def check(node, got):
        lab, kids = node
        if kids:
            assert got.is_internal() and got.label == lab and got.taxon is None
            got_kids = got.child_nodes()
            assert len(got_kids) == len(kids)
            for k, g in zip(kids, got_kids):
                assert g.parent_node is got
                check(k, g)
        else:
            assert got.is_leaf() and got.taxon.label == lab and got.label is None

    for seed in range(40):
        rng = random.Random(seed)
        tree = (None, [build(rng, 6) for _ in range(2)])
        text = write(tree, rng) + ';\n'
        got = newick.Tree.get_from_string(text)
        check(tree, got.seed_node)
        leaves = [l.taxon.label for l in got.leaf_nodes()]
        for name in set(leaves):
            assert got.find_node_with_taxon_label(name).taxon.label == name
