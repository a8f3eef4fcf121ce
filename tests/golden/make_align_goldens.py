"""Goldens for the hmmalign path: run in the build container only (imports the reference read-only).

    python tests/golden/make_align_goldens.py

For four of the fixture's models: a FASTA of full / partial / flanked / tandem homologs and one unrelated sequence
(tests/golden/align/<acc>.unaligned.faa, headers as checkm/hmmerAligner.py:276-287 writes them), the ORACLE's optimal-accuracy
alignment of every sequence formatted as Pfam Stockholm (<acc>.aligned.sto), and the masked FASTA that the REFERENCE's own
HmmerAligner._maskAlignment (checkm/hmmerAligner.py:327-358) makes of it (<acc>.masked.faa) -- the data CheckM consumes.
The GPU test (tests/test_align_gpu.py) runs checkm_b200.HMMERRunner.align on the same FASTA with a one-model HMM file."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, 'align')
CPR = os.path.join(HERE, 'cpr_43_markers.hmm')
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
os.environ.setdefault('CHECKM_DATA_PATH', os.path.join(HERE, 'reduction', 'data'))

import numpy as np  # noqa: E402


def main():
    from oracle import pyoracle as po
    from tools import synth
    from checkm_b200.hmmer import format_alignment
    from checkm.hmmerAligner import HmmerAligner
    hf = po.HmmFile(CPR)
    hm = synth.read_hmms(CPR)
    accs = hf.accs()
    L = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
    os.makedirs(OUT, exist_ok=True)
    for m in (1, 3, 10, 27):
        rng = np.random.default_rng(900 + m)
        h = hm[m]
        bg = lambda n: rng.choice(20, size=n, p=synth.BG).astype(np.uint8)      # noqa: E731
        seqs = [synth.emit_homolog(h, rng, sharpen=0.3),
                np.concatenate([bg(40), synth.emit_homolog(h, rng), bg(25), [27]]),
                synth.emit_homolog(h, rng, k_from=max(1, h.M // 4), k_to=3 * h.M // 4),
                np.concatenate([bg(10), synth.emit_homolog(h, rng, k_from=1, k_to=h.M // 2), bg(60), synth.emit_homolog(h, rng, k_from=h.M // 2 + 1, k_to=h.M)]),
                np.concatenate([synth.emit_homolog(h, rng), bg(8), synth.emit_homolog(h, rng)]),
                np.concatenate([synth.emit_homolog(h, rng, sharpen=0.6), [27]]),
                bg(150)]
        t = seqs[0].copy()
        t[len(t) // 2] = 26
        seqs.append(t)
        seqs = [np.asarray(s, dtype=np.uint8) for s in seqs]
        names = ['bin%d&&c%d_%d' % (i % 3, i + 1, 7 * i + 2) for i in range(len(seqs))]
        descs = ['[e-value=%.4g,score=%.1f]' % (10.0 ** -(5 + i), 50.5 + i) for i in range(len(seqs))]
        base = os.path.join(OUT, accs[m])
        with open(base + '.unaligned.faa', 'w') as f:
            for n, d, s in zip(names, descs, seqs):
                f.write('>%s%s\n%s\n' % (n, (' ' + d) if d else '', ''.join(L[c] for c in s)))
        res = np.concatenate(seqs)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        state = np.concatenate([po.align(hf, m, s)[0] for s in seqs])
        np.save(base + '.state.npy', state)
        with open(base + '.aligned.sto', 'w') as f:
            f.write(format_alignment(names, descs, res, off, state, h.M, 'Pfam', False))
        HmmerAligner(1)._maskAlignment(base + '.aligned.sto', base + '.masked.faa')
        print(accs[m], h.M, [int((st != 0).sum()) for st in np.split(state, off[1:-1])])


if __name__ == '__main__':
    main()
