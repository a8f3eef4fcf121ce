"""Pfam clan / nesting metadata (mirror of checkm/util/pfam.py).

The best-hit-per-clan filter itself (pfam.py:86-147) runs on the device inside ckm_reduce; this class parses
Pfam-A.hmm.dat (pfam.py:34-56), expands marker sets to clan mates (pfam.py:149-168) and turns the metadata into the
integer tables the kernel consumes."""
from collections import defaultdict

from ..common import checkFileExists


class PFAM(object):
    def __init__(self, pfamClanFile):
        self.pfamClanFile = pfamClanFile
        self.idToAcc = {}     # Pfam id -> accession without version
        self.clan = {}        # accession without version -> clan id
        self.nested = {}      # accession without version -> set of nested accessions
        self._parsed = False

    def _readClansAndNesting(self):
        checkFileExists(self.pfamClanFile)
        nestedIds = defaultdict(list)
        curId = None
        curAcc = None
        with open(self.pfamClanFile) as f:
            for line in f:
                if '#=GF ID' in line:
                    curId = line.split()[2].strip()
                elif '#=GF AC' in line:
                    curAcc = line.split()[2].strip()
                    curAcc = curAcc[0:curAcc.rfind('.')]
                    self.idToAcc[curId] = curAcc
                elif '#=GF CL' in line:
                    self.clan[curAcc] = line.split()[2].strip()
                elif '#=GF NE' in line:
                    other = line.split()[2].strip()
                    nestedIds[other].append(curId)
                    nestedIds[curId].append(other)
        for pid, others in nestedIds.items():
            self.nested[self.idToAcc[pid]] = set(self.idToAcc[x] for x in others)
        self._parsed = True

    def ensure_parsed(self):
        if not self._parsed:
            self._readClansAndNesting()

    def pfamIdToClanId(self):
        checkFileExists(self.pfamClanFile)
        d = {}
        acc = None
        with open(self.pfamClanFile) as f:
            for line in f:
                if '#=GF AC' in line:
                    acc = line.split()[2].strip()
                elif '#=GF CL' in line:
                    d[acc] = line.split()[2].strip()
        return d

    def genesInClan(self):
        checkFileExists(self.pfamClanFile)
        d = defaultdict(set)
        acc = None
        with open(self.pfamClanFile) as f:
            for line in f:
                if '#=GF AC' in line:
                    acc = line.split()[2].strip()
                elif '#=GF CL' in line:
                    d[line.split()[2].strip()].add(acc)
        return d

    def genesInSameClan(self, genes):
        """All other genes of the clans spanned by `genes` (versioned accessions)."""
        toClan = self.pfamIdToClanId()
        clans = set(toClan[g] for g in genes if g in toClan)
        members = self.genesInClan()
        everything = set()
        for c in clans:
            everything.update(members[c])
        return everything - genes

    # ---- tables for the device reduction ----
    def reduction_tables(self, accessions):
        """For marker ids `accessions` (list): is_pfam, clan id (-1 = none) and the nesting CSR over list positions."""
        import numpy as np
        self.ensure_parsed()
        n = len(accessions)
        is_pfam = np.zeros(n, dtype=np.uint8)
        clan = np.full(n, -1, dtype=np.int32)
        clan_ids = {}
        stripped = []
        for i, acc in enumerate(accessions):
            is_pfam[i] = 1 if acc.startswith('PF') else 0
            short = acc[0:acc.rfind('.')]          # same slice as pfam.py:112 (drops the last char when there is no '.')
            stripped.append(short)
            c = self.clan.get(short, None)
            if c is not None:
                clan[i] = clan_ids.setdefault(c, len(clan_ids))
        pos = defaultdict(list)
        for i, s in enumerate(stripped):
            pos[s].append(i)
        nest_off = np.zeros(n + 1, dtype=np.int64)
        nest_idx = []
        for i, s in enumerate(stripped):
            if is_pfam[i] and s in self.nested:
                for other in self.nested[s]:
                    nest_idx.extend(pos.get(other, []))
            nest_off[i + 1] = len(nest_idx)
        return is_pfam, clan, nest_off, np.asarray(nest_idx, dtype=np.int32)
