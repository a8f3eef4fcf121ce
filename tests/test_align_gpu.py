"""hmmalign on the GPU (checkm/hmmer.py:76-95 -> `hmmalign`, consumed by checkm/hmmerAligner.py:276-358).

ckm_align = every sequence against ONE model in the configuration hmmalign uses (unihit local, whole sequence as the envelope:
Forward, Backward, posterior decoding, optimal-accuracy fill + traceback).  Bar: the per-residue states (match k / insert k /
flank) equal the oracle's orc_align for every sequence, and the masked FASTA CheckM makes of the alignment file equals the
golden that the REFERENCE's own HmmerAligner._maskAlignment made of the oracle's alignment
(tests/golden/make_align_goldens.py)."""
import os

import numpy as np
import pytest

from conftest import CPR_HMM, GOLDEN
from tools import synth

pytestmark = pytest.mark.gpu
ALI = os.path.join(GOLDEN, 'align')


def _mask(path):
    """checkm/hmmerAligner.py:327-358, restated: residues of the consensus ('x') columns, upper case, '.' -> '-'."""
    seqs, stats, mask = {}, {}, None
    for line in open(path):
        line = line.rstrip()
        if line == '' or line[0] == '#' or line == '//':
            if 'GC RF' in line:
                mask = line.split('GC RF')[1].strip()
            elif '=GS' in line:
                f = line.split()
                stats[f[1]] = f[3].strip()
            continue
        f = line.split()
        seqs[f[0]] = f[1].upper().replace('.', '-').strip()
    out = []
    for sid, seq in seqs.items():
        out.append('>%s %s' % (sid, stats[sid]) if stats else '>' + sid)
        out.append(''.join(seq[i] for i in range(len(seq)) if mask[i] == 'x'))
    return '\n'.join(out) + '\n'


def test_align_states_match_oracle(engine, cpr_models, cpr_oracle, oracle):
    hm = synth.read_hmms(CPR_HMM)
    rng = np.random.default_rng(12)
    for m in (0, 6, 17, 18, 30):                      # short and long models (M = 86 ... 863: several lane-block classes)
        h = hm[m]
        seqs = [synth.emit_homolog(h, rng), np.concatenate([rng.choice(20, size=33, p=synth.BG).astype(np.uint8), synth.emit_homolog(h, rng), [27]]),
                synth.emit_homolog(h, rng, k_from=h.M // 3, k_to=2 * h.M // 3), rng.choice(20, size=90, p=synth.BG).astype(np.uint8),
                np.concatenate([synth.emit_homolog(h, rng), synth.emit_homolog(h, rng)]), np.zeros(0, np.uint8)]
        seqs = [np.asarray(s, np.uint8) for s in seqs]
        off = np.zeros(len(seqs) + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        db = engine.seqdb(np.concatenate(seqs), off)
        state, oasc = engine.align(cpr_models, db, m)
        db.close()
        for i, s in enumerate(seqs):
            exp, sc, rc = oracle.align(cpr_oracle, m, s)
            assert rc == 0 or len(s) == 0
            assert np.array_equal(state[off[i]:off[i + 1]], exp), (m, i)
            assert np.float32(sc) == oasc[i], (m, i, sc, oasc[i])
        assert (state != 0).sum() > 3 * h.M // 2


@pytest.mark.parametrize('acc', ['PF00281.20', 'PF00380.20', 'PF01411.20', 'TIGR01024'])
def test_hmmalign_file_equals_reference_mask(acc, engine, tmp_path):
    from checkm_b200.hmmer import HMMERRunner
    # the per-marker HMM file CheckM fetches before aligning (hmmerAligner.py:411-414): one model
    one = str(tmp_path / 'one.hmm')
    HMMERRunner(mode='fetch').fetch(CPR_HMM, acc, one)
    out = str(tmp_path / (acc + '.aligned.faa'))
    HMMERRunner(mode='align').align(one, os.path.join(ALI, acc + '.unaligned.faa'), out, writeMode='>', outputFormat='Pfam', trim=False)
    assert open(out).read() == open(os.path.join(ALI, acc + '.aligned.sto')).read()
    assert _mask(out) == open(os.path.join(ALI, acc + '.masked.faa')).read()
    # --trim drops the unaligned flanks and nothing else: same masked sequences
    out2 = str(tmp_path / 'trim.sto')
    HMMERRunner(mode='align').align(one, os.path.join(ALI, acc + '.unaligned.faa'), out2, writeMode='>', outputFormat='Pfam', trim=True)
    assert _mask(out2) == _mask(out)
