"""Marker-set model (mirror of checkm/markerSets.py:39-540), with the per-bin HMM extraction done by the engine's
in-memory model database instead of `hmmfetch -f` + `hmmfetch --index` subprocesses (markerSets.py:443-476).

File formats, selection rules and the exclusion list are the reference's; completeness/contamination arithmetic
(`MarkerSet.genomeCheck`, markerSets.py:206-238) is kept here as the host-side statement of what the device
reduction computes (ckm_reduce R4), summed in the same list order."""
import gzip
import logging
import os
import pickle
import shutil
import sys
import tempfile
import uuid

from .defaultValues import DefaultValues
from .hmmerModelParser import HmmModelParser
from .util.pfam import PFAM


class BinMarkerSets(object):
    TAXONOMIC_MARKER_SET = 1
    TREE_MARKER_SET = 2
    HMM_MODELS_SET = 3

    def __init__(self, binId, markerSetType):
        self.logger = logging.getLogger('timestamp')
        self.markerSets = []
        self.binId = binId
        self.markerSetType = markerSetType
        self.selectedLinageSpecificMarkerSet = None

    def numMarkerSets(self):
        return len(self.markerSets)

    def addMarkerSet(self, markerSet):
        self.markerSets.append(markerSet)

    def markerSetIter(self):
        for ms in self.markerSets:
            yield ms

    def getMarkerGenes(self):
        genes = set()
        for ms in self.markerSets:
            genes.update(ms.getMarkerGenes())
        return genes

    def mostSpecificMarkerSet(self):
        return self.markerSets[0]

    def treeMarkerSet(self):
        pass

    def selectedMarkerSet(self):
        if self.markerSetType == self.TAXONOMIC_MARKER_SET:
            return self.mostSpecificMarkerSet()
        if self.markerSetType == self.TREE_MARKER_SET:
            return self.selectedLinageSpecificMarkerSet
        if len(self.markerSets) == 1:
            return self.markerSets[0]
        self.logger.error('Expect a single marker set to be associated with each bin.')
        sys.exit(1)

    def setLineageSpecificSelectedMarkerSet(self, selectedMarkerSetMap):
        """Walk the selected-set map upward until a set carried by this bin is found (reduced-tree hack, :95-121)."""
        selectedId = selectedMarkerSetMap[self.mostSpecificMarkerSet().UID]
        self.selectedLinageSpecificMarkerSet = None
        while self.selectedLinageSpecificMarkerSet is None:
            for ms in self.markerSets:
                if ms.UID == selectedId:
                    self.selectedLinageSpecificMarkerSet = ms
                    break
            if self.selectedLinageSpecificMarkerSet is None:
                selectedId = selectedMarkerSetMap[selectedId]

    def removeMarkers(self, markersToRemove):
        for ms in self.markerSets:
            ms.removeMarkers(markersToRemove)

    def write(self, fout):
        fout.write(self.binId)
        fout.write('\t' + str(len(self.markerSets)))
        for ms in self.markerSets:
            fout.write('\t' + str(ms))
        fout.write('\n')

    def read(self, line):
        """`binId \\t n \\t (uid \\t lineage \\t numGenomes \\t [set([...]), ...]) x n` (markerSets.py:137-154)."""
        fields = line.split('\t')
        for i in range(int(fields[1])):
            uid, lineage, nGenomes, sets = fields[i * 4 + 2:i * 4 + 6]
            self.markerSets.append(MarkerSet(uid, lineage, int(nGenomes), _parse_set_list(sets)))


def _parse_set_list(text):
    """The files hold Python literals such as `[set(['a', 'b']), {'c'}]`; evaluate them without builtins other than set."""
    return eval(text.strip(), {'__builtins__': {}, 'set': set, 'frozenset': frozenset}, {})


class MarkerSet(object):
    def __init__(self, UID, lineageStr, numGenomes, markerSet):
        self.logger = logging.getLogger('timestamp')
        self.UID = UID
        self.lineageStr = lineageStr
        self.numGenomes = numGenomes
        self.markerSet = markerSet

    def __repr__(self):
        return str(self.UID) + '\t' + self.lineageStr + '\t' + str(self.numGenomes) + '\t' + str(self.markerSet)

    def size(self):
        return sum(len(m) for m in self.markerSet), len(self.markerSet)

    def numMarkers(self):
        return self.size()[0]

    def numSets(self):
        return len(self.markerSet)

    def getMarkerGenes(self):
        genes = set()
        for m in self.markerSet:
            genes.update(m)
        return genes

    def removeMarkers(self, markersToRemove):
        kept = []
        for ms in self.markerSet:
            rest = ms - markersToRemove
            if len(rest) != 0:
                kept.append(rest)
        self.markerSet = kept

    def genomeCheck(self, hits, bIndividualMarkers):
        """Completeness / contamination from `{marker: [hits]}` (markerSets.py:206-238)."""
        if bIndividualMarkers:
            present = multi = 0
            for marker in self.getMarkerGenes():
                if marker in hits:
                    present += 1
                    multi += len(hits[marker]) - 1
            return 100 * float(present) / self.numMarkers(), 100 * float(multi) / self.numMarkers()
        comp = cont = 0.0
        for ms in self.markerSet:
            present = multi = 0
            for marker in ms:
                count = len(hits.get(marker, []))
                if count >= 1:
                    present += 1
                    multi += count - 1
            comp += float(present) / len(ms)
            cont += float(multi) / len(ms)
        return 100 * comp / len(self.markerSet), 100 * cont / len(self.markerSet)


_ACC_CACHE = {}


def _hmm_file_accessions(markerFile):
    """Accessions of an HMM file as HmmModelParser.parse() yields them (markerSets.py:265-270), read once per file version:
    the reference re-parses the file on every call, which for a 5,000-model file costs more than searching a bin."""
    key = (os.path.abspath(markerFile), os.path.getmtime(markerFile), os.path.getsize(markerFile))
    accs = _ACC_CACHE.get(key)
    if accs is None:
        accs = [model.acc for model in HmmModelParser(markerFile).parse()]
        _ACC_CACHE.clear()
        _ACC_CACHE[key] = accs
    return accs


class MarkerSetParser(object):
    def __init__(self, threads=1):
        self.logger = logging.getLogger('timestamp')
        self.numThreads = threads
        self._lineage_cache = {}

    # ---- marker sets per bin ----
    def getMarkerSets(self, outDir, binIds, markerFile, excludeMarkersFile=None):
        kind = self.markerFileType(markerFile)
        result = {}
        if kind == BinMarkerSets.TAXONOMIC_MARKER_SET:
            shared = self.parseTaxonomicMarkerSetFile(markerFile)
            for binId in binIds:
                result[binId] = shared
        elif kind == BinMarkerSets.TREE_MARKER_SET:
            result = self.parseLineageMarkerSetFile(markerFile)
        else:
            single = MarkerSet(0, "N/A", -1, [set(_hmm_file_accessions(markerFile))])
            for binId in binIds:
                bms = BinMarkerSets(binId, BinMarkerSets.HMM_MODELS_SET)
                bms.addMarkerSet(single)
                result[binId] = bms
        exclude = set()
        if excludeMarkersFile:
            exclude = self.readExcludeMarkersFile(excludeMarkersFile)
        exclude.update(DefaultValues.MARKERS_TO_EXCLUDE)
        for bms in result.values():
            bms.removeMarkers(exclude)
        return result

    def readExcludeMarkersFile(self, excludeMarkersFile):
        out = set()
        for line in open(excludeMarkersFile):
            if line[0] == '#':
                continue
            out.add(line.strip())
        return out

    def markerFileType(self, markerFile):
        with open(markerFile, 'r') as f:
            header = f.readline()
        if DefaultValues.TAXON_MARKER_FILE_HEADER in header:
            return BinMarkerSets.TAXONOMIC_MARKER_SET
        if DefaultValues.LINEAGE_MARKER_FILE_HEADER in header:
            return BinMarkerSets.TREE_MARKER_SET
        if 'HMMER3' in header:
            return BinMarkerSets.HMM_MODELS_SET
        self.logger.error('Unrecognized file type: ' + markerFile)
        sys.exit(1)

    def parseTaxonomicMarkerSetFile(self, markerSetFile):
        with open(markerSetFile) as f:
            f.readline()
            line = f.readline()
        bms = BinMarkerSets(line.split('\t')[0], BinMarkerSets.TAXONOMIC_MARKER_SET)
        bms.read(line)
        return bms

    def parseLineageMarkerSetFile(self, markerSetFile):
        """One pass over the file and one read of selected_marker_sets.tsv (the reference re-reads that map for
        every line, markerSets.py:498-507; same result)."""
        key = os.path.abspath(markerSetFile)
        stamp = os.path.getmtime(markerSetFile)
        cached = self._lineage_cache.get(key)
        if cached and cached[0] == stamp:
            return cached[1]
        selectedMap = self.parseSelectedMarkerSetMap()
        result = {}
        with open(markerSetFile) as f:
            f.readline()
            for line in f:
                binId = line.split('\t')[0]
                bms = BinMarkerSets(binId, BinMarkerSets.TREE_MARKER_SET)
                bms.read(line)
                bms.setLineageSpecificSelectedMarkerSet(selectedMap)
                result[binId] = bms
        self._lineage_cache[key] = (stamp, result)
        return result

    def parseSelectedMarkerSetMap(self):
        m = {}
        for line in open(DefaultValues.SELECTED_MARKER_SETS):
            fields = line.split('\t')
            m[fields[0]] = fields[1].rstrip()
        return m

    # ---- per-bin HMM selection ----
    def markerAccessions(self, binMarkerSet):
        """Marker genes of all the bin's sets plus every Pfam clan mate (markerSets.py:446-454)."""
        genes = binMarkerSet.getMarkerGenes()
        mates = PFAM(DefaultValues.PFAM_CLAN_FILE).genesInSameClan(genes)
        return genes | mates

    def createHmmModelFile(self, binId, markerFile):
        """Temp HMM file holding the bin's models -- same contract as markerSets.py:326-343 (caller deletes it)."""
        from . import runtime
        kind = self.markerFileType(markerFile)
        out = os.path.join(tempfile.gettempdir(), str(uuid.uuid4()))
        if kind == BinMarkerSets.HMM_MODELS_SET:
            shutil.copyfile(markerFile, out)
            return out
        if kind == BinMarkerSets.TAXONOMIC_MARKER_SET:
            bms = self.parseTaxonomicMarkerSetFile(markerFile)
        else:
            bms = self.parseLineageMarkerSetFile(markerFile)[binId]
        self._createMarkerHMMs(bms, out, bReportProgress=False)
        return out

    def _createMarkerHMMs(self, binMarkerSet, outputFile, bReportProgress=True):
        from . import runtime
        wanted = self.markerAccessions(binMarkerSet)
        if bReportProgress:
            self.logger.info("There are %d genes in the marker set and %d genes from the same PFAM clan." %
                             (len(binMarkerSet.getMarkerGenes()), len(wanted) - len(binMarkerSet.getMarkerGenes())))
        models = runtime.models_for(DefaultValues.HMM_MODELS)
        models.write(models.select(sorted(wanted)), outputFile)

    def createHmmModels(self, outDir, binIds, markerFile):
        kind = self.markerFileType(markerFile)
        result = {}
        if kind == BinMarkerSets.TAXONOMIC_MARKER_SET:
            tmp = self.createHmmModelFile(list(binIds.keys())[0], markerFile)
            models = HmmModelParser(tmp).models()
            os.remove(tmp)
            for binId in binIds:
                result[binId] = models
        elif kind == BinMarkerSets.TREE_MARKER_SET:
            for binId in binIds:
                tmp = self.createHmmModelFile(binId, markerFile)
                result[binId] = HmmModelParser(tmp).models()
                os.remove(tmp)
        else:
            models = HmmModelParser(markerFile).models()
            for binId in binIds:
                result[binId] = models
        return result

    def writeBinModels(self, binIdToModels, filename):
        self.logger.info('Saving HMM info to file.')
        with gzip.open(filename, 'wb') as output:
            pickle.dump(binIdToModels, output, pickle.HIGHEST_PROTOCOL)

    def loadBinModels(self, filename):
        self.logger.info('Reading HMM info from file.')
        with gzip.open(filename, 'rb') as f:
            return pickle.load(f)
