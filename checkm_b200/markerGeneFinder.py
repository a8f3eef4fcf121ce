"""MarkerGeneFinder (mirror of checkm/markerGeneFinder.py:41-181) on the B200 engine.

Same call, same return value (`{binId: {acc: HmmModel}}`), same files under `<outDir>/bins/<binId>/`
(`genes.faa`, the domtblout `tableOut`, the report `hmmerOut`).  What changes is inside: instead of forking
`threads` workers that each spawn hmmfetch + hmmsearch per bin (markerGeneFinder.py:59-89,98-144), the bins of a
batch are digitised once, uploaded as one sequence database and searched in one cascade on this process's GPU with
per-bin query subsets; a CUDA context cannot cross fork(), so no worker processes are created here."""
import gzip
import logging
import os
import shutil
import sys

import numpy as np

from . import runtime
from ._lib import CkmError
from .common import binIdFromFilename, makeSurePathExists
from .defaultValues import DefaultValues
from .hmmer import HMMERRunner, write_domtblout, write_sidecar
from .hmmerModelParser import HmmModel
from .markerSets import BinMarkerSets, MarkerSetParser
from .seqio import read_fasta


def _own_header_keys(info):
    """The header keys HmmModelParser.simpleParse would pick up from this model's own lines."""
    keys = {'name': info.name.decode(), 'leng': int(info.M)}
    if info.acc:
        keys['acc'] = info.acc.decode()
    if info.has_ga:
        keys['ga'] = (info.ga_d[0], info.ga_d[1])
    if info.has_tc:
        keys['tc'] = (info.tc_d[0], info.tc_d[1])
    if info.has_nc:
        keys['nc'] = (info.nc_d[0], info.nc_d[1])
    return keys


def models_as_parsed(infos):
    """`{acc: HmmModel}` for models written in this order to one HMM file and read back by
    HmmModelParser.models(): the reference never clears its header dict between models
    (hmmerModelParser.py:56), so missing ACC/GA/TC/NC keys carry over from the previous model."""
    carry = {'format': 'HMMER3/f'}
    out = {}
    for info in infos:
        carry.update(_own_header_keys(info))
        model = HmmModel(dict(carry))
        out[model.acc] = model
    return out


class MarkerGeneFinder(object):
    def __init__(self, threads):
        self.logger = logging.getLogger('timestamp')
        self.totalThreads = threads
        self.batch_residues = int(os.environ.get('CKM_BATCH_RESIDUES', str(32 * 1024 * 1024)))      # ~32 bins of 3 Mb per search

    def find(self, binFiles, outDir, tableOut, hmmerOut, markerFile, bKeepAlignment, bNucORFs, bCalledGenes):
        HMMERRunner()                       # engine present? (exits like the reference when the tool is missing)
        if not bCalledGenes:
            self._require_prodigal()
        self.logger.info("Identifying marker genes in %d bins on %s:" % (len(binFiles), runtime.engine().device_name()))
        try:
            return self._find(binFiles, outDir, tableOut, hmmerOut, markerFile, bKeepAlignment, bNucORFs, bCalledGenes)
        except CkmError as err:
            self.logger.error('search engine exited with code: %d (%s)' % (err.code, err))
            sys.exit(err.code)

    def _require_prodigal(self):
        try:
            from checkm.prodigal import ProdigalRunner       # gene calling stays CheckM's
            ProdigalRunner('')
        except ImportError:
            self.logger.error("Gene calling (Prodigal) is outside the B200 hot path: supply called genes (--genes).")
            sys.exit(1)

    def _genes_file(self, binFile, binDir, bNucORFs, bCalledGenes):
        if not bCalledGenes:
            from checkm.prodigal import ProdigalRunner
            prodigal = ProdigalRunner(binDir)
            if not prodigal.areORFsCalled(bNucORFs):
                prodigal.run(binFile, bNucORFs)
            return prodigal.aaGeneFile
        saved = os.path.join(binDir, DefaultValues.PRODIGAL_AA)
        if binFile.endswith('.gz'):
            with gzip.open(binFile, 'rt') as fin, open(saved, 'w') as fout:
                shutil.copyfileobj(fin, fout)
        else:
            shutil.copyfile(binFile, saved)
        return saved

    def _find(self, binFiles, outDir, tableOut, hmmerOut, markerFile, bKeepAlignment, bNucORFs, bCalledGenes):
        """Three overlapped stages connected by bounded queues: a reader (gene files -> digitised bins -> batches), the
        searchers (one engine + one host thread each, `CKM_PIPELINE` of them: while one batch is on the GPU the other
        batch's host work proceeds), and a writer (domtblout + side-car per bin).  The reference gets its overlap from
        `threads` forked workers (markerGeneFinder.py:59-89); a CUDA context cannot cross fork(), so these are threads
        of this process -- the library calls release the GIL."""
        import queue
        import threading
        nsearch = max(1, int(os.environ.get('CKM_PIPELINE', '2')))
        engs = runtime.engines(nsearch)
        parser = MarkerSetParser(self.totalThreads)
        kind = parser.markerFileType(markerFile)
        models = runtime.models_for(markerFile if kind == BinMarkerSets.HMM_MODELS_SET else DefaultValues.HMM_MODELS)
        info = models.info()
        all_idx = np.arange(models.n, dtype=np.int32)
        taxon_idx = None
        lineage_sets = None
        if kind == BinMarkerSets.TAXONOMIC_MARKER_SET:
            taxon_idx = models.select(sorted(parser.markerAccessions(parser.parseTaxonomicMarkerSetFile(markerFile))))
        elif kind == BinMarkerSets.TREE_MARKER_SET:
            lineage_sets = parser.parseLineageMarkerSetFile(markerFile)
        # a batch is bounded by residues and by (ORF x HMM) pairs, so that the candidate queues of one search stay small
        max_pairs = int(os.environ.get('CKM_BATCH_PAIRS', str(1 << 32)))
        q_batches, q_out = queue.Queue(maxsize=2 * nsearch), queue.Queue(maxsize=4 * nsearch)
        errors = []
        results = {}
        parsed_cache = {}
        cache_lock = threading.Lock()
        progress = [0]

        def fail(exc):
            errors.append(exc)

        def reader():
            try:
                batch, batch_res, batch_pairs = [], 0, 0
                for binFile in binFiles:
                    if errors:
                        break
                    binId = binIdFromFilename(binFile)
                    binDir = os.path.join(outDir, 'bins', binId)
                    makeSurePathExists(binDir)
                    genes = self._genes_file(binFile, binDir, bNucORFs, bCalledGenes)
                    names, descs, residues, offsets = read_fasta(genes)
                    if kind == BinMarkerSets.HMM_MODELS_SET:
                        midx = all_idx
                    elif kind == BinMarkerSets.TAXONOMIC_MARKER_SET:
                        midx = taxon_idx
                    else:
                        midx = models.select(sorted(parser.markerAccessions(lineage_sets[binId])))
                    pairs = len(names) * len(midx)
                    if batch and (batch_res + len(residues) > self.batch_residues or batch_pairs + pairs > max_pairs):
                        q_batches.put(batch)
                        batch, batch_res, batch_pairs = [], 0, 0
                    batch.append((binId, binDir, names, descs, residues, offsets, midx))
                    batch_res += len(residues)
                    batch_pairs += pairs
                if batch and not errors:
                    q_batches.put(batch)
            except BaseException as exc:          # noqa: B902 -- surfaced on the calling thread
                fail(exc)
            finally:
                for _ in range(nsearch):
                    q_batches.put(None)

        def searcher(eng):
            try:
                while True:
                    batch = q_batches.get()
                    if batch is None:
                        break
                    if errors:
                        continue
                    res = np.concatenate([b[4] for b in batch])
                    lens = np.concatenate([np.diff(b[5]) for b in batch])
                    off = np.zeros(len(lens) + 1, dtype=np.int64)
                    off[1:] = np.cumsum(lens)
                    binof = np.repeat(np.arange(len(batch), dtype=np.int32), [len(b[2]) for b in batch])
                    sdb = eng.seqdb(res, off, binof, len(batch))
                    try:
                        if kind == BinMarkerSets.TREE_MARKER_SET:
                            midx = np.concatenate([b[6] for b in batch]).astype(np.int32)
                            boff = np.zeros(len(batch) + 1, dtype=np.int64)
                            boff[1:] = np.cumsum([len(b[6]) for b in batch])
                            hits = eng.search(models, sdb, model_idx=midx, E=0.1, domE=0.1, bin_model_offsets=boff)
                        else:
                            hits = eng.search(models, sdb, model_idx=batch[0][6], E=0.1, domE=0.1)
                    finally:
                        sdb.close()
                    # rows come back grouped by bin: hand every bin its slice
                    bounds = np.searchsorted(hits['bin'], np.arange(len(batch) + 1))
                    seq_base = 0
                    for i, b in enumerate(batch):
                        q_out.put((b, hits[bounds[i]:bounds[i + 1]], i, seq_base))
                        seq_base += len(b[2])
            except BaseException as exc:          # noqa: B902
                fail(exc)
            finally:
                q_out.put(None)

        def writer():
            done_searchers = 0
            try:
                while done_searchers < nsearch:
                    item = q_out.get()
                    if item is None:
                        done_searchers += 1
                        continue
                    if errors:
                        continue
                    (binId, binDir, names, descs, _r, _o, midx), sub, bin_index, seq_base = item
                    table = os.path.join(binDir, tableOut)
                    write_domtblout(models, sub, bin_index, seq_base, names, descs, table)
                    side = sub.copy()
                    side['seq'] -= seq_base
                    side['bin'] = 0
                    write_sidecar(table, side, names, descs, models)
                    if bKeepAlignment:
                        with open(os.path.join(binDir, hmmerOut), 'w') as f:
                            f.write('# checkm_b200: alignment display is not produced (SURVEY.md 8f2); %d domtblout rows\n' % len(sub))
                    key = midx.tobytes()
                    with cache_lock:
                        if key not in parsed_cache:
                            parsed_cache[key] = models_as_parsed([info[int(m)] for m in midx])
                        results[binId] = parsed_cache[key]
                    progress[0] += 1
                    if self.logger.getEffectiveLevel() <= logging.INFO:
                        sys.stderr.write('    Finished processing %d of %d (%.2f%%) bins.\r' % (progress[0], len(binFiles), progress[0] * 100.0 / len(binFiles)))
                        sys.stderr.flush()
            except BaseException as exc:          # noqa: B902
                fail(exc)
                while done_searchers < nsearch:   # keep draining so that the searchers can finish
                    if q_out.get() is None:
                        done_searchers += 1

        threads = [threading.Thread(target=reader, name='ckm-reader')] + \
                  [threading.Thread(target=searcher, args=(e,), name='ckm-search') for e in engs] + \
                  [threading.Thread(target=writer, name='ckm-writer')]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        if self.logger.getEffectiveLevel() <= logging.INFO:
            sys.stderr.write('\n')
        return {binIdFromFilename(f): results[binIdFromFilename(f)] for f in binFiles}
