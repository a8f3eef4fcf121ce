"""Lineage-specific marker-set selection on the placed genome tree.

Mirror of the part of `checkm/treeParser.py` that stands between the phylogenetic search (`hmmer.tree.txt`, reduced by
`ResultsParser`) and the per-bin marker file the main search consumes: `TreeParser.getBinMarkerSets`
(treeParser.py:468-553) with its helpers `_getMarkerSet` (:344-380), `_findDomainNode` (:223-258), `_getNextNamedNode`
(:327-342), `_removeInvalidLineageMarkerGenes` (:441-466), `_readLineageSpecificGenesToRemove` (:430-439),
`readNodeMetadata` (:555-584), and the three small tree look-ups `readLineageMetadata` (:586-629), `getBinTaxonomy`
(:182-221), `getInsertionBranchId` (:151-180).  pplacer itself stays external: the input is the Newick file it leaves at
`<outDir>/storage/tree/concatenated.tre`.

What differs from the reference is cost, not results:
  * the tree is read by `util/newick.py` (no dendropy) and bins are found through a label index, not a scan per bin;
  * `missing_duplicate_genes_50.tsv` is parsed once per TreeParser (the reference re-reads and re-`eval`s it per bin);
  * "does this clade hold a reference genome" (`_findDomainNode`'s leaf scan per ancestor) is one bottom-up pass;
  * the marker-set literal of a node is parsed once and shared between the bins that pass through it.
Reference behaviour kept on purpose: `_getMarkerSet` builds the set from the LAST LABELLED node it looked at, so a walk
that reaches the root without a node meeting the criteria returns the root's set under the lineage name 'root'; bins are
written in `os.listdir` order; a `PF` accession is matched against the removal list as `pfamNNNNN` without version.
Inside a full CheckM install the class derives from CheckM's own `TreeParser`, so `tree_qa`'s reports keep working and call
into the methods below; `_findDomainNode` therefore uses only the node API both tree classes share."""
import logging
import os
import sys

from .defaultValues import DefaultValues
from .markerSets import MarkerSet, BinMarkerSets, _parse_set_list
from .common import checkDirExists, getBinIdsFromOutDir
from .util import newick

try:                                              # inside a full CheckM install (dendropy present): CheckM's reports
    from checkm.treeParser import TreeParser as _ReportBase          # (printSummary, reportBinTaxonomy, ...) stay available
except Exception:                                 # stand-alone: the selection and the look-ups below are all there is
    _ReportBase = object


class TreeParser(_ReportBase):
    def __init__(self):
        self.logger = logging.getLogger('timestamp')
        self.lineageSpecificGenesToRemove = None
        self._set_cache = {}

    # ---- inputs -------------------------------------------------------------------------------------------------
    def _readTree(self, outDir):
        treeFile = os.path.join(outDir, 'storage', 'tree', DefaultValues.PPLACER_TREE_OUT)
        return newick.Tree.get_from_path(treeFile, schema='newick')

    def readNodeMetadata(self):
        """`genome_tree.metadata.tsv`: one line per internal node of the reference tree (treeParser.py:555-584)."""
        uniqueIdToLineageStatistics = {}
        metadataFile = os.path.join(DefaultValues.GENOME_TREE_DIR, DefaultValues.GENOME_TREE_METADATA)
        with open(metadataFile) as f:
            f.readline()
            for line in f:
                lineSplit = line.rstrip().split('\t')
                d = {}
                d['# genomes'] = int(lineSplit[1])
                d['taxonomy'] = lineSplit[2]
                try:
                    d['bootstrap'] = float(lineSplit[3])
                except Exception:
                    d['bootstrap'] = 'NA'
                d['gc mean'] = float(lineSplit[4])
                d['gc std'] = float(lineSplit[5])
                d['genome size mean'] = float(lineSplit[6]) / 1e6
                d['genome size std'] = float(lineSplit[7]) / 1e6
                d['gene count mean'] = float(lineSplit[8])
                d['gene count std'] = float(lineSplit[9])
                d['marker set'] = lineSplit[10].rstrip()
                uniqueIdToLineageStatistics[lineSplit[0]] = d
        return uniqueIdToLineageStatistics

    def _readLineageSpecificGenesToRemove(self):
        """uid -> genes lost or duplicated in that lineage (treeParser.py:430-439); read once."""
        if self.lineageSpecificGenesToRemove is not None:
            return
        table = {}
        with open(os.path.join(DefaultValues.GENOME_TREE_DIR, DefaultValues.GENOME_TREE_MISSING_DUPLICATE)) as f:
            for line in f:
                lineSplit = line.split('\t')
                table[lineSplit[0]] = _parse_set_list(lineSplit[1]).union(_parse_set_list(lineSplit[2]))
        self.lineageSpecificGenesToRemove = table

    def _markerSetOf(self, stats):
        text = stats['marker set']
        sets = self._set_cache.get(text)
        if sets is None:
            sets = self._set_cache[text] = _parse_set_list(text)
        return [set(s) for s in sets]          # every MarkerSet owns its sets (callers remove markers in place)

    # ---- walks ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _firstLabelledAncestor(node):
        parentNode = node.parent_node
        while parentNode is not None:
            if parentNode.label:
                return parentNode
            parentNode = parentNode.parent_node
        return None

    @staticmethod
    def _referenceCladeFlags(tree):
        """id(node) -> True when the clade holds a reference genome (a leaf named IMG_*); one bottom-up pass."""
        order = list(tree.preorder_node_iter())
        flags = {}
        for node in reversed(order):
            if node.is_leaf():
                flags[id(node)] = bool(node.taxon is not None and node.taxon.label.startswith('IMG_'))
            else:
                flags[id(node)] = any(flags[id(c)] for c in node._children)
        return flags

    def _findDomainNode(self, binNode, flags=None):
        """First labelled internal node below the first ancestor that holds a reference genome (treeParser.py:223-258)."""
        curNode = binNode.parent_node
        while True:
            if curNode is None:
                self.logger.error('Failed to associate bin with a domain. Please report this bug.')
                sys.exit(1)
            if flags is not None:
                found = flags[id(curNode)]
            else:
                found = any(leaf.taxon.label.startswith('IMG_') for leaf in curNode.leaf_nodes())
            if found:
                break
            curNode = curNode.parent_node
        queue = [curNode]
        head = 0
        while head < len(queue):
            curNode = queue[head]
            head += 1
            if curNode.label:
                return curNode
            for child in curNode.child_nodes():
                if child.is_internal():
                    queue.append(child)
        self.logger.error('Failed to associate bin with a domain. Please report this bug.')
        sys.exit(1)

    def _getNextNamedNode(self, node, uniqueIdToLineageStatistics):
        parentNode = node.parent_node
        while parentNode is not None:
            if parentNode.label:
                trustedStats = uniqueIdToLineageStatistics[parentNode.label.split('|')[0]]
                if trustedStats['taxonomy'] != '':
                    return trustedStats['taxonomy']
            parentNode = parentNode.parent_node
        return 'root'

    def _getMarkerSet(self, parentNode, tree, uniqueIdToLineageStatistics, numGenomesMarkers, bootstrap, bForceDomain,
                      bRequireTaxonomy):
        """Marker set of the first node at or above `parentNode` that meets the selection criteria (treeParser.py:344-380)."""
        selectedParentNode = parentNode
        taxonomyStr = 'root'
        trustedUniqueId = stats = None
        while True:
            if selectedParentNode.label:                               # pplacer-inserted nodes carry no label
                tokens = selectedParentNode.label.split('|')
                trustedUniqueId = tokens[0]
                nodeTaxonomy = tokens[1]
                stats = uniqueIdToLineageStatistics[trustedUniqueId]
                if ((stats['# genomes'] == 'NA' or int(stats['# genomes']) >= numGenomesMarkers)
                        and (stats['bootstrap'] == 'NA' or int(stats['bootstrap']) >= bootstrap)):
                    if not bForceDomain or nodeTaxonomy in ('k__Bacteria', 'k__Archaea'):
                        if not bRequireTaxonomy or stats['taxonomy'] != '':
                            taxonomyStr = stats['taxonomy']
                            if not bRequireTaxonomy and stats['taxonomy'] == '':
                                taxonomyStr = self._getNextNamedNode(selectedParentNode, uniqueIdToLineageStatistics)
                            break
            if selectedParentNode.parent_node is None:
                break
            selectedParentNode = selectedParentNode.parent_node
        if stats is None:
            self.logger.error('No labelled node between the insertion point and the root of the genome tree.')
            sys.exit(1)
        taxonomyStr = taxonomyStr.split(';')[-1]
        markerSet = MarkerSet(trustedUniqueId, taxonomyStr, int(stats['# genomes']), self._markerSetOf(stats))
        return selectedParentNode, markerSet

    def _removeInvalidLineageMarkerGenes(self, markerSet, lineageSpecificMarkersToRemove):
        """Drop genes subject to lineage-specific loss / duplication; co-location stays the trusted set's (:441-466)."""
        finalMarkerSet = []
        for ms in markerSet.markerSet:
            s = set()
            for gene in ms:
                geneIdToTest = gene
                if geneIdToTest.startswith('PF'):
                    geneIdToTest = gene.replace('PF', 'pfam')
                    geneIdToTest = geneIdToTest[0:geneIdToTest.rfind('.')]
                if geneIdToTest not in lineageSpecificMarkersToRemove:
                    s.add(gene)
            if s:
                finalMarkerSet.append(s)
        return MarkerSet(markerSet.UID, markerSet.lineageStr, markerSet.numGenomes, finalMarkerSet)

    # ---- the entry point -------------------------------------------------------------------------------------------
    def getBinMarkerSets(self, outDir, markerFile, numGenomesMarkers, bootstrap, bNoLineageSpecificRefinement, bForceDomain,
                         bRequireTaxonomy, resultsParser, minUnique, maxMulti):
        """Write the lineage marker file: per bin every marker set met on the way from its insertion point to the root."""
        self.logger.info('Determining marker sets for each genome bin.')
        binIds = getBinIdsFromOutDir(outDir)
        uniqueIdToLineageStatistics = self.readNodeMetadata()
        tree = self._readTree(outDir)
        rootNode = tree.seed_node
        flags = None
        with open(markerFile, 'w') as fout:
            fout.write(DefaultValues.LINEAGE_MARKER_FILE_HEADER + '\n')
            for binId in binIds:
                node = tree.find_node_with_taxon_label(binId)
                binMarkerSets = BinMarkerSets(binId, BinMarkerSets.TREE_MARKER_SET)
                if node is None:                                       # bin is not in the tree: the root's set
                    node, markerSet = self._getMarkerSet(rootNode, tree, uniqueIdToLineageStatistics, numGenomesMarkers,
                                                         bootstrap, bForceDomain, bRequireTaxonomy)
                    binMarkerSets.addMarkerSet(markerSet)
                else:
                    parentNode = self._firstLabelledAncestor(node)
                    if parentNode is None:
                        self.logger.error('Failed to find lineage-specific statistics for inserted bin: ' + binId)
                        sys.exit(1)
                    if parentNode.parent_node is None:
                        # inserted on the bacterial or archaeal branch below the root: start under the domain node so that
                        # the domain's own marker set is on the path
                        if flags is None:
                            flags = self._referenceCladeFlags(tree)
                        curNode = self._findDomainNode(node, flags).child_nodes()[0]
                    else:
                        curNode = node
                    lineageSpecificRefinement = None
                    if not bNoLineageSpecificRefinement:
                        self._readLineageSpecificGenesToRemove()
                        uniqueId = parentNode.label.split('|')[0]
                        if uniqueId not in self.lineageSpecificGenesToRemove:
                            self.logger.error('No lineage-specific gene list for node %s (%s).' % (
                                uniqueId, DefaultValues.GENOME_TREE_MISSING_DUPLICATE))
                            sys.exit(1)
                        lineageSpecificRefinement = self.lineageSpecificGenesToRemove[uniqueId]
                    uniqueHits, multiCopyHits = resultsParser.results[binId].countUniqueHits()
                    tempForceDomain = bForceDomain or (uniqueHits < minUnique) or (multiCopyHits > maxMulti)
                    while curNode.parent_node is not None:
                        curNode, markerSet = self._getMarkerSet(curNode.parent_node, tree, uniqueIdToLineageStatistics,
                                                                numGenomesMarkers, bootstrap, tempForceDomain, bRequireTaxonomy)
                        if not bNoLineageSpecificRefinement:
                            markerSet = self._removeInvalidLineageMarkerGenes(markerSet, lineageSpecificRefinement)
                        binMarkerSets.addMarkerSet(markerSet)
                binMarkerSets.write(fout)

    # ---- look-ups used by tree_qa -----------------------------------------------------------------------------------
    def readLineageMetadata(self, outDir, binIds):
        """Statistics of the first labelled ancestor of every bin (treeParser.py:586-629)."""
        uniqueIdToLineageStatistics = self.readNodeMetadata()
        tree = self._readTree(outDir)
        binIdToLineageStatistics = {}
        for binId in binIds:
            node = tree.find_node_with_taxon_label(binId)
            if node is None:
                d = {k: 'NA' for k in ('# genomes', 'gc mean', 'gc std', 'genome size mean', 'genome size std',
                                       'gene count mean', 'gene count std', 'marker set')}
                d['taxonomy'] = 'unresolved'
                binIdToLineageStatistics[binId] = d
                continue
            parentNode = self._firstLabelledAncestor(node)
            if parentNode is None:
                self.logger.error('Failed to find lineage-specific statistics for inserted bin: ' + node.taxon.label)
                sys.exit(1)
            binIdToLineageStatistics[binId] = uniqueIdToLineageStatistics[parentNode.label.split('|')[0]]
        return binIdToLineageStatistics

    def getInsertionBranchId(self, outDir, binIds):
        checkDirExists(outDir)
        checkDirExists(os.path.join(outDir, 'storage', 'tree'))
        tree = self._readTree(outDir)
        binIdToUID = {}
        for binId in binIds:
            node = tree.find_node_with_taxon_label(binId)
            if node is None:
                binIdToUID[binId] = 'NA'
                continue
            parentNode = self._firstLabelledAncestor(node)
            if parentNode is None:
                self.logger.error('Failed to find lineage-specific statistics for inserted bin: ' + binId)
                sys.exit(1)
            binIdToUID[binId] = parentNode.label.split('|')[0]
        return binIdToUID

    def getBinTaxonomy(self, outDir, binIds):
        """Taxon strings of all labelled ancestors, most general first (treeParser.py:182-221)."""
        checkDirExists(outDir)
        checkDirExists(os.path.join(outDir, 'storage', 'tree'))
        tree = self._readTree(outDir)
        flags = None
        binIdToTaxonomy = {}
        for binId in binIds:
            node = tree.find_node_with_taxon_label(binId)
            if node is None:
                binIdToTaxonomy[binId] = 'NA'
                continue
            taxaStr = None
            parentNode = node.parent_node
            while parentNode is not None:
                if parentNode.label:
                    tokens = parentNode.label.split('|')
                    if tokens[1] != '':
                        taxaStr = tokens[1] + ';' + taxaStr if taxaStr else tokens[1]
                parentNode = parentNode.parent_node
            if not taxaStr:
                if flags is None:
                    flags = self._referenceCladeFlags(tree)
                taxaStr = self._findDomainNode(node, flags).label.split('|')[1] + ' (root)'
            binIdToTaxonomy[node.taxon.label] = taxaStr
        return binIdToTaxonomy
