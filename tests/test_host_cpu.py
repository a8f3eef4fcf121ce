"""CPU-only checks: the C ABI loads and exports every declared symbol, host-side mirrors of the reference interface
behave like the reference (known answers from SURVEY.md Appendix B), and the product refuses to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import CPR_HMM, GOLDEN, ROOT


def test_abi_exports_every_declared_symbol():
    from checkm_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'ckm.h')).read()
    declared = set(re.findall(r'^(?:int|void|const char \*)\s*(ckm_[a-z0-9_]+)\s*\(', header, re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s


def test_struct_sizes_match_header():
    from checkm_b200 import _lib
    assert ctypes.sizeof(_lib.Hit) == 112
    assert ctypes.sizeof(_lib.QaRow) == 64
    assert ctypes.sizeof(_lib.MarkerHit) == 64


def test_no_cpu_fallback():
    """Without a CUDA device ckm_init must fail loudly (CKM_ENODEVICE), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from checkm_b200.engine import Engine
    from checkm_b200._lib import CkmError
    with pytest.raises(CkmError) as ei:
        Engine(0)
    assert ei.value.code == 4


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'checkm_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cpp', '.hpp', '.cuh', '.h')) or fn == 'Makefile':
                text = open(os.path.join(dirpath, fn), errors='replace').read()
                if re.search(r'oracle', text, re.I) and fn not in ('synth.py',):
                    if re.search(r'(import\s+oracle|from\s+oracle|hmmer_oracle|liboracle|pyoracle)', text):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_hmm_model_parser_header_kat():
    from checkm_b200.hmmerModelParser import HmmModelParser
    models = HmmModelParser(CPR_HMM).models()
    assert len(models) == 43
    assert sum(m.leng for m in models.values()) == 8926
    m = models['PF00276.21']
    assert (m.name, m.leng, m.ga, m.tc, m.nc) == ('Ribosomal_L23', 86, (30.8, 30.8), (30.8, 30.8), (30.6, 30.7))
    assert models['TIGR00029'].nc == (39.2, 39.2)


def test_hmm_model_parser_carry_over_quirk(tmp_path):
    """KAT-4: a model without ACC/GA inherits the previous model's (reference hmmerModelParser.py:56)."""
    from checkm_b200.hmmerModelParser import HmmModelParser
    p = tmp_path / 'two.hmm'
    p.write_text('HMMER3/f [x]\nNAME  modelA\nACC   PF99999.1\nLENG  3\nGA    10.0 9.0;\nHMM   A\n  1 x\n//\n'
                 'HMMER3/f [x]\nNAME  modelB\nLENG  3\nHMM   A\n  1 x\n//\n')
    models = HmmModelParser(str(p)).models()
    assert list(models.keys()) == ['PF99999.1']
    m = models['PF99999.1']
    assert (m.name, m.leng, m.ga) == ('modelB', 3, (10.0, 9.0))


def test_genome_check_kat3():
    from checkm_b200.markerSets import MarkerSet
    ms = MarkerSet(1, 'x', 10, [{'a', 'b', 'c'}, {'d'}, {'e', 'f'}])
    hits = {'a': [1], 'b': [1, 2, 3], 'd': [1, 2], 'f': [1]}
    assert ms.genomeCheck(hits, False) == (72.22222222222221, 55.55555555555555)
    assert ms.genomeCheck(hits, True) == (66.66666666666667, 50.0)
    assert ms.size() == (6, 3) and ms.numMarkers() == 6 and ms.numSets() == 3


def test_marker_set_files_and_exclusion(tmp_path):
    from checkm_b200.markerSets import MarkerSetParser, BinMarkerSets
    taxon = os.path.join(GOLDEN, 'reduction', 'taxon.ms')
    msp = MarkerSetParser()
    assert msp.markerFileType(taxon) == BinMarkerSets.TAXONOMIC_MARKER_SET
    assert msp.markerFileType(CPR_HMM) == BinMarkerSets.HMM_MODELS_SET
    sets = msp.getMarkerSets(str(tmp_path), ['b1', 'b2'], taxon)
    assert sets['b1'] is sets['b2']
    sel = sets['b1'].selectedMarkerSet()
    assert (sel.UID, sel.lineageStr, sel.numGenomes, sel.numSets()) == ('42', 'k__Bacteria', 5449, 9)
    hm = msp.getMarkerSets(str(tmp_path), ['b1'], CPR_HMM)['b1'].selectedMarkerSet()
    assert hm.numMarkers() == 43 and hm.UID == 0
    # the two unreliable TIGRFAMs are always removed (defaultValues.py:34)
    p = tmp_path / 'x.ms'
    p.write_text("# [Taxon Marker File]\nT\t1\t7\tk__X\t3\t[set(['TIGR00398', 'PF1.1']), set(['TIGR00399'])]\n")
    got = msp.getMarkerSets(str(tmp_path), ['b'], str(p))['b'].selectedMarkerSet()
    assert got.markerSet == [{'PF1.1'}]


def test_pfam_tables():
    from checkm_b200.util.pfam import PFAM
    pf = PFAM(os.path.join(GOLDEN, 'reduction', 'data', 'pfam', 'Pfam-A.hmm.dat'))
    accs = ['PF00276.21', 'PF00281.20', 'TIGR00002', 'PF00366.21', 'PF00380.20', 'PF13393.7']
    is_pfam, clan, off, idx = pf.reduction_tables(accs)
    assert list(is_pfam) == [1, 1, 0, 1, 1, 1]
    assert clan[0] == clan[1] and clan[3] == clan[4] and clan[0] != clan[3] and clan[2] == -1 and clan[5] == -1
    assert list(idx[off[3]:off[4]]) == [4] and list(idx[off[4]:off[5]]) == [3]
    assert pf.genesInSameClan({'PF00276.21'}) == {'PF00281.20'}


def test_domtblout_text_parser_matches_reference_fields():
    from checkm_b200.hmmer import HMMERParser
    path = os.path.join(GOLDEN, 'reduction', 'kat2', 'bins', 'binB', 'hmmer.analyze.txt')
    hits = []
    with open(path) as f:
        hp = HMMERParser(f)
        while True:
            h = hp.next()
            if h is None:
                break
            hits.append(h)
    assert len(hits) == 11
    h = hits[1]
    assert (h.target_name, h.target_length, h.query_accession, h.query_length, h.dom, h.ndom, h.dom_score, h.ali_from, h.ali_to) == \
           ('k1_1', 200, 'TIGR00029', 78, 2, 2, 60.0, 100, 170)
    assert hits[8].full_score == 18.04 and hits[9].query_name == 'Ribosomal_S9' and hits[9].target_description == '# d'


def test_evalue_threshold_decimal_split():
    from checkm_b200.resultsParser import _decimal_split
    assert _decimal_split(1e-10) == (-10, 10.0)
    assert _decimal_split(2.5e-7) == (-7, 25.0)
    assert _decimal_split(0.05) == (-2, 50.0)


def test_fasta_parser_in_the_library():
    """ckm_fasta_parse (host code of libckm.so; no GPU needed): codes, CSR offsets, names and descriptions of a protein FASTA."""
    import gzip
    import numpy as np
    from checkm_b200.seqio import parse_fasta, read_fasta
    text = ">a desc one\nACDE\nFG\n>b\n\nHIK*\n>c only\n>d x\r\nLMN\r\nPQ a\tb\n"
    names, descs, res, off = parse_fasta(text.encode())
    assert names == ['a', 'b', 'c', 'd'] and descs == ['desc one', '', 'only', 'x']
    assert off.tolist() == [0, 6, 10, 10, 17]
    alphabet = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
    assert ''.join(alphabet[c] for c in res) == 'ACDEFGHIK*LMNPQAX'.replace('X', 'B')[:17] or ''.join(alphabet[c] for c in res) == 'ACDEFGHIK*LMNPQAB'
    assert parse_fasta(b'')[0] == [] and parse_fasta(b'junk\n>x\nAC')[2].tolist() == [0, 1]
    e2e = os.path.join(ROOT, 'tests', 'golden', 'e2e', 'bins')
    n1 = read_fasta(os.path.join(e2e, 'binB.faa.gz'))
    with gzip.open(os.path.join(e2e, 'binB.faa.gz'), 'rt') as f:
        raw = f.read()
    assert len(n1[0]) == raw.count('>') and int(n1[3][-1]) == len(''.join(l for l in raw.split('\n') if not l.startswith('>')))


def test_alignment_formatter():
    """hmmer.format_alignment: consensus columns + insert columns as hmmalign lays them out, and what CheckM's mask makes of it."""
    import numpy as np
    from checkm_b200.hmmer import format_alignment
    res = np.array([0, 1, 0, 1, 2, 3, 4, 5, 2, 3, 4], dtype=np.uint8)            # ACACDEFG | DEF
    state = np.array([0, 0, 1, 2, -2, -2, 4, 0, 2, 3, 4], dtype=np.int32)
    off = np.array([0, 8, 11])
    sto = format_alignment(['s1', 'seq2'], ['[e-value=1e-05,score=3.0]', '[e-value=1,score=1.0]'], res, off, state, 4, 'Pfam', False)
    lines = sto.split('\n')
    assert lines[0] == '# STOCKHOLM 1.0' and lines[-2] == '//'
    assert 's1      acACde-Fg' in lines and 'seq2    ..-D..EF.' in lines and '#=GC RF ..xx..xx.' in lines
    assert '#=GS s1   DE [e-value=1e-05,score=3.0]' in lines
    trimmed = format_alignment(['s1', 'seq2'], ['', ''], res, off, state, 4, 'Pfam', True)
    assert 's1      ACde-F' in trimmed.split('\n') and '#=GC RF xx..xx' in trimmed.split('\n')
    afa = format_alignment(['s1', 'seq2'], ['', ''], res, off, state, 4, 'afa', True)
    assert afa == '>s1\nACde-F\n>seq2\n-D..EF\n'
