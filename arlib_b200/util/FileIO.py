"""Space-separated triple files -- mirror of the reference's util/FileIO.py:5-31."""
import os


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
        rows = []
        with open(file) as fh:
            for line in fh:
                f = line.strip().split(' ')
                rows.append([f[0], f[1], float(f[2])])
        return rows
