"""TEST + BENCH INFRASTRUCTURE (not product). Seeded synthetic workloads for tests and bench.py.

Not part of the search path: it only manufactures inputs -- protein bins in Prodigal naming
(`>c<j>_<k> # start # end # strand # ID=j_k`, sequences ending in '*') whose ORFs are background
residues with marker homologs planted by sampling the profile HMMs themselves.
"""
import numpy as np

AMINO = "ACDEFGHIKLMNPQRSTVWY"
ALPHABET = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"

# Swiss-Prot 50.8 background
BG = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092,
               0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364,
               0.0540687, 0.0673417, 0.0114135, 0.0304133])
BG = BG / BG.sum()


class PyHmm:
    """Emission/transition probabilities of one HMMER3/f model, for sampling only."""

    def __init__(self, name, acc, M, mat, ins, t):
        self.name, self.acc, self.M, self.mat, self.ins, self.t = name, acc, M, mat, ins, t


def _p(tok):
    return 0.0 if tok == '*' else float(np.exp(-float(tok)))


def read_hmms(path):
    """Minimal HMMER3/f reader (body included) used only to sample homologs."""
    models = []
    with open(path) as f:
        lines = f.read().split('\n')
    i, n = 0, len(lines)
    while i < n:
        if not lines[i].startswith('HMMER3'):
            i += 1
            continue
        name = acc = None
        M = 0
        while not lines[i].startswith('HMM '):
            tag = lines[i].split(None, 1)
            if tag and tag[0] == 'NAME':
                name = tag[1].strip()
            elif tag and tag[0] == 'ACC':
                acc = tag[1].strip()
            elif tag and tag[0] == 'LENG':
                M = int(tag[1])
            i += 1
        i += 2
        mat = np.zeros((M + 1, 20))
        ins = np.zeros((M + 1, 20))
        t = np.zeros((M + 1, 7))
        for k in range(M + 1):
            tok = lines[i].split()
            if k == 0:
                if tok[0] == 'COMPO':
                    i += 1
                    tok = lines[i].split()
                ins[0] = [_p(x) for x in tok[:20]]
            else:
                mat[k] = [_p(x) for x in tok[1:21]]
                i += 1
                ins[k] = [_p(x) for x in lines[i].split()[:20]]
            i += 1
            t[k] = [_p(x) for x in lines[i].split()[:7]]
            i += 1
        models.append(PyHmm(name, acc if acc else name, M, mat, ins, t))
        i += 1
    return models


def emit_homolog(hmm, rng, k_from=1, k_to=None, sharpen=0.0):
    """Sample one path through match/insert/delete states k_from..k_to of the core model; returns residue codes.
    sharpen > 0: a match state emits its consensus residue with that probability (closer homologs; the default 0 draws
    nothing extra, so seeded bins made without it are unchanged)."""
    k_to = hmm.M if k_to is None else k_to
    out = []
    k = k_from
    state = 'M'
    while True:
        if state == 'M':
            p = hmm.mat[k] / hmm.mat[k].sum()
            if sharpen > 0.0 and rng.random() < sharpen:
                out.append(int(np.argmax(p)))
            else:
                out.append(rng.choice(20, p=p))
        elif state == 'I':
            p = hmm.ins[k] / hmm.ins[k].sum()
            out.append(rng.choice(20, p=p))
        if k >= k_to and state != 'I':
            break
        tr = hmm.t[k]
        if state == 'M':
            pr = np.array([tr[0], tr[1], tr[2]])
            nxt = 'MID'[rng.choice(3, p=pr / pr.sum())]
        elif state == 'I':
            pr = np.array([tr[3], tr[4]])
            nxt = 'MI'[rng.choice(2, p=pr / pr.sum())]
            if k >= k_to and nxt == 'M':
                break
        else:
            pr = np.array([tr[5], tr[6]])
            nxt = 'MD'[rng.choice(2, p=pr / pr.sum())]
        if nxt != 'I':
            k += 1
            if k > k_to:
                break
        state = nxt
    return np.array(out, dtype=np.uint8)


def random_lengths(rng, n, mean=310, lo=30, hi=3000):
    """Shifted-gamma ORF lengths."""
    x = rng.gamma(shape=2.2, scale=(mean - lo) / 2.2, size=n) + lo
    return np.clip(x, lo, hi).astype(np.int64)


class Bin:
    """One synthetic bin: concatenated digitised residues, offsets, names, descriptions, and the planted truth."""

    def __init__(self, bin_id, residues, offsets, names, descs, planted):
        self.bin_id, self.residues, self.offsets = bin_id, residues, offsets
        self.names, self.descs, self.planted = names, descs, planted

    @property
    def nseq(self):
        return len(self.offsets) - 1

    def seq(self, i):
        return self.residues[self.offsets[i]:self.offsets[i + 1]]

    def fasta(self):
        out = []
        for i in range(self.nseq):
            s = ''.join(ALPHABET[c] for c in self.seq(i))
            out.append('>%s %s\n%s\n' % (self.names[i], self.descs[i], s))
        return ''.join(out)


def make_bin(bin_id, hmms, seed, n_orfs=1900, mean_len=310, copies=(0, 1, 1, 1, 2), split_prob=0.08,
             orfs_per_contig=40, degenerate_prob=0.002, tandem_prob=0.0, max_len=3000, sharpen=0.0):
    """Build one bin: background ORFs, 0-2 planted copies of every family, some split
    across adjacent ORFs (k, k+1 of one contig) to exercise the adjacent-ORF merge, a sprinkle of X/B/Z."""
    rng = np.random.default_rng(seed)
    lens = random_lengths(rng, n_orfs, mean=mean_len, hi=max_len)
    seqs = [rng.choice(20, size=int(l), p=BG).astype(np.uint8) for l in lens]
    planted = []
    free = list(rng.permutation(n_orfs - 1))
    used = set()

    def take():
        while free:
            o = int(free.pop())
            if o not in used and (o + 1) not in used and (o % orfs_per_contig) != orfs_per_contig - 1:
                used.add(o)
                used.add(o + 1)
                return o
        raise RuntimeError("not enough ORFs to plant into")

    for mi, h in enumerate(hmms):
        for _ in range(int(rng.choice(copies))):
            o = take()
            if rng.random() < split_prob and h.M >= 60:
                cut = int(rng.integers(h.M // 3, 2 * h.M // 3))
                a = emit_homolog(h, rng, 1, cut, sharpen)
                b = emit_homolog(h, rng, cut + 1, h.M, sharpen)
                for orf, frag in ((o, a), (o + 1, b)):
                    lf, rf = int(rng.integers(5, 60)), int(rng.integers(5, 60))
                    seqs[orf] = np.concatenate([rng.choice(20, size=lf, p=BG), frag, rng.choice(20, size=rf, p=BG)]).astype(np.uint8)
                planted.append((mi, o, 'split'))
            else:
                frag = emit_homolog(h, rng, sharpen=sharpen)
                if rng.random() < tandem_prob:
                    frag = np.concatenate([frag, rng.choice(20, size=int(rng.integers(3, 15)), p=BG), emit_homolog(h, rng, sharpen=sharpen)])
                lf, rf = int(rng.integers(0, 120)), int(rng.integers(0, 120))
                seqs[o] = np.concatenate([rng.choice(20, size=lf, p=BG), frag, rng.choice(20, size=rf, p=BG)]).astype(np.uint8)
                planted.append((mi, o, 'full'))
    # rare degenerate symbols
    for s in seqs:
        m = rng.random(len(s)) < degenerate_prob
        if m.any():
            s[m] = rng.choice([26, 21, 23], size=int(m.sum()))
    names, descs, chunks = [], [], []
    pos = 1
    for i, s in enumerate(seqs):
        contig, k = i // orfs_per_contig + 1, i % orfs_per_contig + 1
        if k == 1:
            pos = 1
        nt = 3 * (len(s) + 1)
        names.append('c%d_%d' % (contig, k))
        descs.append('# %d # %d # %d # ID=%d_%d;partial=00;start_type=ATG;rbs_motif=None;rbs_spacer=None;gc_cont=0.500'
                     % (pos, pos + nt - 1, 1 if rng.random() < 0.5 else -1, contig, k))
        pos += nt + int(rng.integers(10, 200))
        chunks.append(np.concatenate([s, np.array([27], dtype=np.uint8)]))   # trailing '*'
    offsets = np.zeros(len(chunks) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(c) for c in chunks])
    return Bin(bin_id, np.concatenate(chunks), offsets, names, descs, planted)


def perturbed_model_lengths(rng, n, lo=30, hi=1500, mean=240):
    """Model lengths for the ~5k-HMM stand-in of configs #3-#5 (log-normal around the Pfam/TIGRFAM mean)."""
    x = rng.lognormal(mean=np.log(mean) - 0.18, sigma=0.6, size=n)
    return np.clip(x, lo, hi).astype(np.int64)


def fitted_stats(M):
    """(MSV mu, Viterbi mu, Forward tau, lambda) as a function of model length, least-squares fitted to the 43 real,
    HMMER-calibrated CPR models (residual s.d. 0.17 / 0.16 / 0.24 bits) -- keeps the filter pass rates of synthetic
    models near their nominal 2% / 0.1% / 1e-5."""
    x = np.log2(max(M, 2))
    return (-2.599826 - 1.03052763 * x, -2.57804831 - 1.13368047 * x, 1.65869172 - 0.86205053 * x, 0.69315 + 1.8 / max(M, 2))


def _nl(p):
    return '      *' if p <= 0.0 else '%9.5f' % (-np.log(p))[:9].strip().rjust(9) if False else ('%.5f' % (-np.log(p)))


def write_hmms(path, models, stats=None):
    """Write PyHmm-like models as HMMER3/f text.  `stats[i]` = (msv_mu, vit_mu, fwd_tau, lambda) or None."""
    with open(path, 'w') as f:
        for i, h in enumerate(models):
            M = h.M
            mu = stats[i] if stats is not None else fitted_stats(M)
            f.write('HMMER3/f [3.1b2 | February 2015]\n')
            f.write('NAME  %s\n' % h.name)
            if h.acc and h.acc != h.name:
                f.write('ACC   %s\n' % h.acc)
            f.write('LENG  %d\nALPH  amino\nRF    no\nMM    no\nCONS  no\nCS    no\nMAP   no\n' % M)
            f.write('NSEQ  10\nEFFN  1.000000\nCKSUM 0\n')
            ga = getattr(h, 'ga', None)
            if ga is not None:
                f.write('GA    %.2f %.2f;\n' % (ga, ga))
            f.write('STATS LOCAL MSV      %8.4f  %.5f\n' % (mu[0], mu[3]))
            f.write('STATS LOCAL VITERBI  %8.4f  %.5f\n' % (mu[1], mu[3]))
            f.write('STATS LOCAL FORWARD  %8.4f  %.5f\n' % (mu[2], mu[3]))
            f.write('HMM          ' + '        '.join(AMINO) + '   \n')
            f.write('            m->m     m->i     m->d     i->m     i->i     d->m     d->d\n')
            compo = h.mat[1:].mean(axis=0)
            f.write('  COMPO   ' + '  '.join(_nl(p) for p in compo) + '\n')
            for k in range(M + 1):
                if k > 0:
                    f.write('%7d   ' % k + '  '.join(_nl(p) for p in h.mat[k]) + '      %d - - - -\n' % k)
                f.write('          ' + '  '.join(_nl(p) for p in h.ins[k]) + '\n')
                f.write('          ' + '  '.join(('      *' if p <= 0 else _nl(p)) for p in h.t[k]) + '\n')
            f.write('//\n')


def resample_model(src_models, M, rng, name, acc=None):
    """A model of length M stitched from random windows of real models (keeps realistic emission/transition rows)."""
    mat = np.zeros((M + 1, 20))
    ins = np.zeros((M + 1, 20))
    t = np.zeros((M + 1, 7))
    k = 1
    first = src_models[int(rng.integers(len(src_models)))]
    ins[0] = first.ins[0]
    t[0] = first.t[0]
    while k <= M:
        s = src_models[int(rng.integers(len(src_models)))]
        w = int(min(M - k + 1, rng.integers(10, 60), s.M - 2))
        a = int(rng.integers(1, s.M - w))
        mat[k:k + w] = s.mat[a:a + w]
        ins[k:k + w] = s.ins[a:a + w]
        t[k:k + w] = s.t[a:a + w]
        k += w
    # last node: no insert/delete exits (HMMER convention: MM=1? file stores m->m, *, m->d=*; i->m, i->i; d->m=1, d->d=*)
    last = src_models[0]
    t[M] = last.t[last.M]
    ins[M] = last.ins[last.M]
    h = PyHmm(name, acc if acc else name, M, mat, ins, t)
    return h


def make_model_db(path, src_path, lengths, seed=0, prefix='SYN'):
    """Writes a synthetic HMM database with the given model lengths (configs #3-#5 stand-in)."""
    rng = np.random.default_rng(seed)
    src = read_hmms(src_path)
    out = []
    for i, M in enumerate(lengths):
        out.append(resample_model(src, int(M), rng, '%s%05d' % (prefix, i), 'PF%05d.1' % (90000 + i) if i % 2 == 0 else 'TIGR%05d' % (90000 + i)))
    write_hmms(path, out)
    return out


def make_model_db_fast(path, src_path, lengths, seed=0, prefix='SYN'):
    """Like make_model_db but stitches the source file's own text lines (no number formatting): fast enough to
    write the ~5,000-model stand-in of configs #3-#5 in seconds.  Returns the list of model lengths written."""
    rng = np.random.default_rng(seed)
    with open(src_path) as f:
        text = f.read().split('\n')
    # index the source: per model, header STATS lines and the three text lines of every node
    src = []
    i = 0
    while i < len(text):
        if not text[i].startswith('HMMER3'):
            i += 1
            continue
        M = 0
        stats = []
        while not text[i].startswith('HMM '):
            if text[i].startswith('LENG'):
                M = int(text[i].split()[1])
            elif text[i].startswith('STATS'):
                stats.append(text[i])
            i += 1
        hdr2 = text[i:i + 2]
        i += 2
        compo = None
        if text[i].lstrip().startswith('COMPO'):
            compo = text[i]
            i += 1
        node0 = text[i:i + 2]
        i += 2
        nodes = []
        for k in range(1, M + 1):
            m_line = text[i]
            nodes.append((m_line[m_line.index(str(k)) + len(str(k)):][:181], text[i + 1], text[i + 2]))
            i += 3
        src.append(dict(M=M, stats=stats, hdr2=hdr2, compo=compo, node0=node0, nodes=nodes))
        i += 1
    out = []
    with open(path, 'w') as f:
        for n, M in enumerate(lengths):
            M = int(M)
            first = src[int(rng.integers(len(src)))]
            f.write('HMMER3/f [3.1b2 | February 2015]\nNAME  %s%05d\n' % (prefix, n))
            f.write('ACC   %s\n' % ('PF%05d.1' % (90000 + n) if n % 2 == 0 else 'TIGR%05d' % (90000 + n)))
            f.write('LENG  %d\nALPH  amino\nRF    no\nMM    no\nCONS  no\nCS    no\nMAP   no\nNSEQ  10\nEFFN  1.000000\nCKSUM 0\n' % M)
            st = fitted_stats(M)
            f.write('STATS LOCAL MSV      %8.4f  %.5f\nSTATS LOCAL VITERBI  %8.4f  %.5f\nSTATS LOCAL FORWARD  %8.4f  %.5f\n' % (st[0], st[3], st[1], st[3], st[2], st[3]))
            f.write(first['hdr2'][0] + '\n' + first['hdr2'][1] + '\n')
            if first['compo'] is not None:
                f.write(first['compo'] + '\n')
            f.write(first['node0'][0] + '\n' + first['node0'][1] + '\n')
            k = 1
            buf = []
            while k <= M:
                s = src[int(rng.integers(len(src)))]
                w = int(min(M - k + 1, rng.integers(10, 60), s['M'] - 2))
                a = int(rng.integers(0, s['M'] - w - 1))
                for z in range(w):
                    ml, il, tl = s['nodes'][a + z]
                    if k + z == M:
                        tl = src[0]['nodes'][-1][2]          # a proper last node: no exits into insert/delete
                    buf.append('%7d%s\n%s\n%s\n' % (k + z, ml, il, tl))
                k += w
            f.write(''.join(buf))
            f.write('//\n')
            out.append(M)
    return out
