"""Generates the reduction goldens by running the REFERENCE's own code (checkm.resultsParser / markerSets / util.pfam,
imported read-only from /root/reference) on domtblout inputs.  Run in the build container only:

    python tests/golden/make_reduction_goldens.py

Inputs and expected outputs are written under tests/golden/reduction/ and committed; the GPU tests replay them
through checkm_b200 (tests/test_reduction_gpu.py).  Cases:
  kat1, kat2   the hand-written known-answer tables of SURVEY.md Appendix B
  synth0..2    domtblout written by the CPU oracle for seeded synthetic bins (split genes -> adjacent-ORF merges)
  stress       random rows built to hit every rule: clan overlap / nesting / clan-less Pfams, best-domain replacement,
               adjacency chains, non-integer ORF suffixes, thresholds on the .1 boundary, E-value ties
"""
import io
import json
import os
import shutil
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, 'reduction')
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
DATA = os.path.join(OUT, 'data')
os.makedirs(os.path.join(DATA, 'pfam'), exist_ok=True)
os.environ['CHECKM_DATA_PATH'] = DATA

import numpy as np   # noqa: E402

CPR = os.path.join(HERE, 'cpr_43_markers.hmm')

CLAN_FILE = """# STOCKHOLM 1.0
#=GF ID   Ribosomal_L23
#=GF AC   PF00276.21
#=GF CL   CL0001
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_L5
#=GF AC   PF00281.20
#=GF CL   CL0001
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S17
#=GF AC   PF00366.21
#=GF CL   CL0002
#=GF NE   Ribosomal_S9
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S9
#=GF AC   PF00380.20
#=GF CL   CL0002
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S11
#=GF AC   PF00411.20
#=GF CL   CL0002
//
"""

KAT1 = """contig1_1 - 100 Ribosomal_L23 PF00276.21 86 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 86  5 95  5 95 0.98 # 1 # 300 # 1 # ID=1_1
contig1_1 - 100 Ribosomal_L5  PF00281.20 57 1e-20  70.0 0.1 1 1 1e-23 1e-20 69.0 0.1  1 57 10 80 10 80 0.98 # 1 # 300 # 1 # ID=1_1
contig1_2 - 100 TIGR00002     TIGR00002  78 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40  5 75  5 75 0.98 # 301 # 600 # 1 # ID=1_2
contig1_3 - 100 TIGR00002     TIGR00002  78 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1 41 78  5 75  5 75 0.98 # 601 # 900 # 1 # ID=1_3
contig2_7 - 100 TIGR00029     TIGR00029  78 1e-30  10.0 0.1 1 1 1e-33 1e-30  9.0 0.1 41 78  5 75  5 75 0.98 # 601 # 900 # 1 # ID=2_7
"""

KAT2 = """k1_1 - 200 TIGR00029 TIGR00029  78 1e-30 100.0 0.1 1 2 1e-33 1e-30 50.0 0.1  1 40   5  75   5  75 0.98 # d
k1_1 - 200 TIGR00029 TIGR00029  78 1e-30 100.0 0.1 2 2 1e-33 1e-30 60.0 0.1 41 78 100 170 100 170 0.98 # d
k1_2 - 100 TIGR00029 TIGR00029  78 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40   5  75   5  75 0.98 # d
k1_3 - 100 TIGR00029 TIGR00029  78 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40   5  75   5  75 0.98 # d
k2_x - 100 TIGR00060 TIGR00060  78 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40   5  75   5  75 0.98 # d
k2_y - 100 TIGR00060 TIGR00060  78 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40   5  75   5  75 0.98 # d
k3_5 - 100 TIGR00061 TIGR00061 100 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40   5  34   5  75 0.98 # d
k3_6 - 100 TIGR00062 TIGR00062 100 1e-30 100.0 0.1 1 1 1e-33 1e-30 99.0 0.1  1 40   5  35   5  75 0.98 # d
k4_1 - 100 PF13393.7 PF13393.7 100 1e-30 18.04 0.1 1 1 1e-33 1e-30 18.0 0.1  1 40   5  95   5  75 0.98 # d
k4_2 - 100 Ribosomal_S9 PF00380.20 121 1e-30 22.1 0.1 1 1 1e-33 1e-30 22.1 0.1 1 40  5  95   5  75 0.98 # d
k4_3 - 100 Ribosomal_S9 PF00380.20 121 1e-30 22.1 0.1 1 1 1e-33 1e-30 22.0 0.1 1 40  5  95   5  75 0.98 # d
"""


def stress_table(models, seed):
    """Random domtblout rows, grouped by query like hmmsearch writes them."""
    rng = np.random.default_rng(seed)
    accs = list(models.keys())
    lines = []
    for acc in accs:
        m = models[acc]
        thr = (m.nc if ('TIGR' in acc and m.nc) else (m.ga or m.tc or m.nc))[0]
        n = int(rng.choice([0, 0, 1, 1, 2, 3, 5, 8]))
        rows = []
        for _ in range(n):
            contig = int(rng.integers(1, 5))
            style = rng.random()
            if style < 0.8:
                orf = 'c%d_%d' % (contig, int(rng.integers(1, 12)))
            elif style < 0.9:
                orf = 'c%d_x%d' % (contig, int(rng.integers(1, 4)))        # non-integer suffix
            else:
                orf = 'plainname%d' % int(rng.integers(1, 4))             # no underscore at all
            ndom = int(rng.choice([1, 1, 1, 2, 3]))
            fullsc = thr + float(rng.choice([-0.1, 0.0, 0.04, 0.05, 0.06, 3.3, 20.0, 75.5]))
            fe = float(rng.choice([1e-30, 1.2e-25, 1.25e-25, 3e-12, 9.9e-11, 1e-10, 1.1e-10]))
            for d in range(ndom):
                domsc = fullsc - float(rng.choice([0.0, 0.0, 0.1, 0.5, 2.0, 30.0]))
                hf = int(rng.integers(1, m.leng // 2 + 1))
                ht = int(rng.integers(hf, m.leng + 1))
                af = int(rng.integers(1, 150))
                at = af + int(rng.choice([5, int(0.29 * m.leng), int(0.3 * m.leng) + 1, int(0.7 * m.leng), m.leng]))
                ie = fe * float(rng.choice([1.0, 1.0, 10.0, 0.1]))
                rows.append((fe, '%-12s - %5d %-16s %-12s %5d %9.2g %6.1f %5.1f %3d %3d %9.2g %9.2g %6.1f %5.1f %5d %5d %5d %5d %5d %5d %4.2f # %d # %d # 1 # ID=x'
                             % (orf, at + 40, m.name, acc, m.leng, fe, fullsc, 0.1, d + 1, ndom, ie / 10, ie, domsc, 0.1, hf, ht, af, at,
                                max(1, af - 2), at + 2, 0.9, af, at)))
        rows.sort(key=lambda r: r[0])
        lines.extend(r[1] for r in rows)
    return '\n'.join(lines) + '\n'


def dump_hits(rm):
    out = []
    for acc, hits in rm.markerHits.items():
        out.append([acc, [[h.target_name, h.target_length, h.hmm_from, h.hmm_to, h.ali_from, h.ali_to, h.env_from, h.env_to,
                           h.dom_score, h.full_score, h.full_e_value, h.i_evalue] for h in hits]])
    return out


def main():
    from checkm.hmmerModelParser import HmmModelParser
    from checkm.markerSets import MarkerSetParser, MarkerSet, BinMarkerSets
    from checkm.resultsParser import ResultsParser
    from checkm.defaultValues import DefaultValues

    with open(os.path.join(DATA, 'pfam', 'Pfam-A.hmm.dat'), 'w') as f:
        f.write(CLAN_FILE)
    models = HmmModelParser(CPR).models()

    cases = {'kat1': {'binA': KAT1}, 'kat2': {'binB': KAT2}}
    # oracle-written tables for synthetic bins
    from oracle import pyoracle as po
    from tools import synth
    hf = po.HmmFile(CPR)
    hm = synth.read_hmms(CPR)
    synth_tables = {}
    for i in range(3):
        b = synth.make_bin('syn%d' % i, hm, seed=100 + i, n_orfs=260, split_prob=0.5, max_len=900, tandem_prob=0.1)
        rp = po.search(hf, b.residues, b.offsets, nthreads=8)
        path = os.path.join(OUT, '_tmp_dom.txt')
        po.write_domtblout(rp, hf, b.names, b.descs, path)
        po.free_results(rp)
        synth_tables['syn%d' % i] = open(path).read()
        os.remove(path)
    cases['synth'] = synth_tables
    cases['stress'] = {'st%d' % s: stress_table(models, 7 + s) for s in range(6)}

    # a collocated marker-set file over the 43 accessions (taxon format, markerSets.py:137-154)
    accs = sorted(models.keys())
    groups = [accs[i:i + 5] for i in range(0, len(accs), 5)]
    taxon_line = 'Bacteria\t1\t42\tk__Bacteria\t5449\t' + str([set(g) for g in groups])
    with open(os.path.join(OUT, 'taxon.ms'), 'w') as f:
        f.write('# [Taxon Marker File]\n' + taxon_line + '\n')

    golden = {}
    for case, tables in cases.items():
        cdir = os.path.join(OUT, case)
        shutil.rmtree(cdir, ignore_errors=True)
        os.makedirs(os.path.join(cdir, 'storage'))
        with open(os.path.join(cdir, 'storage', 'bin_stats.analyze.tsv'), 'w') as f:
            for binId in tables:
                f.write("%s\t{'GC': 0.5, 'Genome size': 1000}\n" % binId)
        for binId, text in tables.items():
            os.makedirs(os.path.join(cdir, 'bins', binId))
            with open(os.path.join(cdir, 'bins', binId, 'hmmer.analyze.txt'), 'w') as f:
                f.write('# target name accession tlen query name accession qlen ...\n' + text + '#\n# [ok]\n')
        binIds = list(tables.keys())
        g = {}
        for label, kwargs in (('default', {}), ('noadj', {'bSkipAdjCorrection': True}), ('nopseudo', {'bSkipPseudoGeneCorrection': True}),
                              ('ignore', {'bIgnoreThresholds': True})):
            rp = ResultsParser({b: models for b in binIds})
            rp.analyseResults(cdir, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt', **kwargs)
            entry = {}
            for which, mfile in (('hmm', CPR), ('taxon', os.path.join(OUT, 'taxon.ms'))):
                ms = MarkerSetParser().getMarkerSets(cdir, binIds, mfile)
                per_bin = {}
                for b in binIds:
                    rm = rp.results[b]
                    per_bin[b] = {'counts_colloc': rm.geneCountsForSelectedMarkerSet(ms[b], False),
                                  'counts_indiv': rm.geneCountsForSelectedMarkerSet(ms[b], True),
                                  'unique': list(rm.countUniqueHits())}
                buf = io.StringIO()

                class _AAI:
                    aaiMeanBinHetero = {}
                with redirect_stdout(buf):
                    rp.printSummary(1, _AAI(), ms, False, None, True, '', None)
                entry[which] = {'bins': per_bin, 'table1': buf.getvalue()}
                for fmt in (5, 6, 8):
                    buf = io.StringIO()
                    with redirect_stdout(buf):
                        rp.printSummary(fmt, _AAI(), ms, False, None, True, '', None)
                    entry[which]['table%d' % fmt] = buf.getvalue()
            entry['hits'] = {b: dump_hits(rp.results[b]) for b in binIds}
            g[label] = entry
        golden[case] = g
    with open(os.path.join(OUT, 'reduction_goldens.json'), 'w') as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    shutil.rmtree(os.path.join(OUT, '_tmp'), ignore_errors=True)
    print('cases', {k: list(v.keys()) for k, v in cases.items()})


if __name__ == '__main__':
    main()
