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
        self.batch_residues = int(os.environ.get('CKM_BATCH_RESIDUES', str(256 * 1024 * 1024)))

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
        eng = runtime.engine()
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

        binIdToModels = {}
        parsed_cache = {}
        batch = []                           # (binId, binDir, names, descs, residues, offsets, model_idx)
        batch_res = 0
        done = 0

        def flush():
            nonlocal batch, batch_res, done
            if not batch:
                return
            res = np.concatenate([b[4] for b in batch]) if batch else np.zeros(0, np.uint8)
            lens = np.concatenate([np.diff(b[5]) for b in batch])
            off = np.zeros(len(lens) + 1, dtype=np.int64)
            off[1:] = np.cumsum(lens)
            binof = np.concatenate([np.full(len(b[2]), i, dtype=np.int32) for i, b in enumerate(batch)])
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
            seq_base = 0
            for i, (binId, binDir, names, descs, _r, _o, midx) in enumerate(batch):
                table = os.path.join(binDir, tableOut)
                sub = hits[hits['bin'] == i]
                write_domtblout(models, sub, i, seq_base, names, descs, table)
                side = sub.copy()
                side['seq'] -= seq_base
                side['bin'] = 0
                write_sidecar(table, side, names, descs, models)
                if bKeepAlignment:
                    with open(os.path.join(binDir, hmmerOut), 'w') as f:
                        f.write('# checkm_b200: alignment display is not produced (SURVEY.md 8f2); %d domtblout rows\n' % len(sub))
                key = midx.tobytes()
                if key not in parsed_cache:
                    parsed_cache[key] = models_as_parsed([info[int(m)] for m in midx])
                binIdToModels[binId] = parsed_cache[key]
                seq_base += len(names)
                done += 1
                if self.logger.getEffectiveLevel() <= logging.INFO:
                    sys.stderr.write('    Finished processing %d of %d (%.2f%%) bins.\r' % (done, len(binFiles), done * 100.0 / len(binFiles)))
                    sys.stderr.flush()
            batch = []
            batch_res = 0

        for binFile in binFiles:
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
            batch.append((binId, binDir, names, descs, residues, offsets, midx))
            batch_res += len(residues)
            if batch_res >= self.batch_residues:
                flush()
        flush()
        if self.logger.getEffectiveLevel() <= logging.INFO:
            sys.stderr.write('\n')
        return binIdToModels
