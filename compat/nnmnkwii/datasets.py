"""nnmnkwii.datasets shim (host-side data pipeline used by reference train.py:71-136,705-716)."""
import numpy as np


class FileDataSource(object):
    """Base class: subclasses implement collect_files() and collect_features(path)."""

    def collect_files(self):
        raise NotImplementedError

    def collect_features(self, *args):
        raise NotImplementedError


class FileSourceDataset(object):
    """Lazily loaded list of per-utterance feature arrays."""

    def __init__(self, file_data_source):
        self.file_data_source = file_data_source
        collected = self.file_data_source.collect_files()
        self.multiple = isinstance(collected, tuple)
        self.collected_files = np.asarray(collected).T if self.multiple else np.atleast_2d(collected).T

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self[i] for i in range(*idx.indices(len(self)))]
        paths = self.collected_files[idx]
        return self.file_data_source.collect_features(*paths)

    def __len__(self):
        return len(self.collected_files)


class MemoryCacheDataset(object):
    """LRU-less memo cache in front of a dataset (cache_size entries)."""

    def __init__(self, dataset, cache_size=777):
        self.dataset, self.cache_size, self.cache = dataset, cache_size, {}

    def __getitem__(self, idx):
        if idx not in self.cache:
            if len(self.cache) >= self.cache_size:
                self.cache.pop(next(iter(self.cache)))
            self.cache[idx] = self.dataset[idx]
        return self.cache[idx]

    def __len__(self):
        return len(self.dataset)
