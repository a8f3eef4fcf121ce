"""HMM header model (mirror of checkm/hmmerModelParser.py:27-83).

`HmmModel` carries name / acc / leng / ga / tc / nc exactly as the reference's objects do -- they are pickled
into storage/checkm_hmm_info.pkl.gz and must keep unpickling.  The reference parser never resets its header
dictionary between models (hmmerModelParser.py:56), so a model lacking ACC/GA/TC/NC inherits the previous
model's; `simpleParse` reproduces that."""


class HmmModelError(Exception):
    pass


class _HmmModel(object):
    """Plain attribute bag: ga/tc/nc default to None, acc defaults to name when no ACC key was seen."""

    def __init__(self, keys):
        self.ga = None
        self.tc = None
        self.nc = None
        if 'acc' not in keys:
            self.acc = keys['name']
        for key, value in keys.items():
            setattr(self, key, value)


try:                                    # drop-in inside a CheckM install: use CheckM's own class, so that the pickles
    from checkm.hmmerModelParser import HmmModel      # (storage/checkm_hmm_info.pkl.gz, markerSets.py:524-540) are interchangeable
except Exception:                       # stand-alone: same attributes, our module path
    HmmModel = _HmmModel
    HmmModel.__name__ = 'HmmModel'
    HmmModel.__qualname__ = 'HmmModel'


def _cutoff_pair(text):
    parts = text.split()
    if len(parts) != 2:
        raise HmmModelError
    return (float(parts[0].replace(';', '')), float(parts[1].replace(';', '')))


class HmmModelParser(object):
    def __init__(self, hmmFile):
        self.hmmFile = open(hmmFile)

    def models(self):
        out = {}
        for model in self.simpleParse():
            out[model.acc] = model
        return out

    def simpleParse(self):
        header = dict()                      # deliberately shared across models, like the reference
        for line in self.hmmFile:
            if line.startswith('HMMER'):
                header['format'] = line.rstrip()
            elif line.startswith('HMM'):
                for line in self.hmmFile:    # skip the body up to the record terminator
                    if line.startswith('//'):
                        yield HmmModel(header)
                        break
            else:
                fields = line.rstrip().split(None, 1)
                if len(fields) != 2:
                    raise HmmModelError
                tag, value = fields
                if tag in ('ACC', 'NAME'):
                    header[tag.lower()] = value
                elif tag == 'LENG':
                    header['leng'] = int(value)
                elif tag in ('GA', 'TC', 'NC'):
                    header[tag.lower()] = _cutoff_pair(value)

    def parse(self):
        """Full header parse (hmmerModelParser.py:85-129): every tag kept, a few typed."""
        header = dict()
        for line in self.hmmFile:
            if line.startswith('HMMER'):
                header['format'] = line.rstrip()
            elif line.startswith('HMM'):
                for line in self.hmmFile:
                    if line.startswith('//'):
                        yield HmmModel(header)
                        break
            else:
                fields = line.rstrip().split(None, 1)
                if len(fields) != 2:
                    raise HmmModelError
                tag, value = fields
                key = tag.lower()
                if tag in ('LENG', 'NSEQ', 'CKSUM'):
                    header[key] = int(value)
                elif tag in ('RF', 'CS', 'MAP'):
                    header[key] = value.lower() != 'no'
                elif tag == 'EFFN':
                    header[key] = float(value)
                elif tag in ('GA', 'TC', 'NC'):
                    header[key] = _cutoff_pair(value)
                elif tag == 'STATS':
                    params = value.split()
                    if params[0] != 'LOCAL' or params[1] not in ('MSV', 'VITERBI', 'FORWARD'):
                        raise HmmModelError
                    header[('stats_' + params[0] + '_' + params[1]).lower()] = (float(params[2]), float(params[3]))
                else:
                    header[key] = value
