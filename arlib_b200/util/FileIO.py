"""Space-separated triple files -- mirror of the reference's util/FileIO.py:5-31.

``load_data_set`` returns the reference's list of ``[user, item, float(weight)]`` rows.  The rows come from a C parser
(pandas) when the file is regular, and the list then also carries the three parsed columns (``TripleRows.columns``) so
that DataLoader can index a million rows with array operations instead of per-row dict updates (every
attack step writes the poisoned train.txt as text and re-parses it before retraining)."""
import contextlib
import gc
import os

import numpy as np


@contextlib.contextmanager
def no_gc():
    """Bulk construction of millions of small containers: the generational collector re-scans them again and again
    (3-4x the build time at 1 M rows) and can free nothing -- switch it off for the duration."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class TripleRows(list):
    """list of [user, item, weight] rows + the columns they were built from (object, object, float64 arrays).
    ``columns`` describes the list only while nobody resized it: consumers check the length."""
    columns = None

    def parsed_columns(self):
        c = self.columns
        return c if c is not None and len(c[0]) == len(self) else None


def rows_from_columns(users, items, weights):
    with no_gc():
        rows = TripleRows(map(list, zip(users.tolist(), items.tolist(), weights.tolist())))
    rows.columns = (users, items, weights)
    return rows


class FileIO(object):
    @staticmethod
    def write_file(dir, file, content, op='w'):
        """util/FileIO.py:9-14"""
        os.makedirs(dir, exist_ok=True)
        with open(dir + file, op) as fh:
            fh.writelines(content)

    @staticmethod
    def delete_file(file_path):
        """util/FileIO.py:16-19"""
        if os.path.exists(file_path):
            os.remove(file_path)

    @staticmethod
    def load_data_set(file):
        """util/FileIO.py:21-31 -- '<user> <item> <weight>' per line -> [str, str, float]."""
        try:
            import pandas as pd
            df = pd.read_csv(file, sep=' ', header=None, usecols=[0, 1, 2], dtype={0: str, 1: str, 2: np.float64},
                             engine='c', keep_default_na=False, na_filter=False, skip_blank_lines=False,
                             quoting=3, skipinitialspace=False,
                             float_precision='round_trip')      # == float(text) of the reference, to the last ulp
            return rows_from_columns(df[0].to_numpy(dtype=object), df[1].to_numpy(dtype=object),
                                     df[2].to_numpy(dtype=np.float64))
        except Exception:
            pass                       # irregular file (ragged lines, stray blanks ...): the reference's own loop
        rows = []
        with open(file) as fh:
            for line in fh:
                f = line.strip().split(' ')
                rows.append([f[0], f[1], float(f[2])])
        return rows
