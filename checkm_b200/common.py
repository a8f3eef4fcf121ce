"""Small path helpers used by the hot path (same behaviour as checkm/common.py)."""
import errno
import logging
import os
import sys


def checkFileExists(inputFile):
    if not os.path.exists(inputFile):
        logging.getLogger('timestamp').error('Input file does not exists: ' + inputFile + '\n')
        sys.exit(1)


def makeSurePathExists(path):
    if not path:
        return
    try:
        os.makedirs(path)
    except OSError as exc:
        if exc.errno != errno.EEXIST:
            logging.getLogger('timestamp').error('Specified path does not exist: ' + path + '\n')
            sys.exit(1)


def binIdFromFilename(filename):
    """Bin id = file name without directory and without its last extension (a trailing .gz is dropped first)."""
    binId = os.path.basename(filename)
    if binId.endswith('.gz'):
        binId = binId[0:-3]
    return os.path.splitext(binId)[0]


def reassignStdOut(outFile):
    oldStdOut = sys.stdout
    if outFile != '':
        try:
            sys.stdout = open(outFile, 'w')
        except IOError:
            logging.getLogger('timestamp').error('Error diverting stdout to file: ' + outFile)
            sys.exit(1)
    return oldStdOut


def restoreStdOut(outFile, oldStdOut):
    if outFile != '':
        try:
            sys.stdout.close()
            sys.stdout = oldStdOut
        except IOError:
            logging.getLogger('timestamp').error('Error restoring stdout: ' + outFile)
            sys.exit(1)
