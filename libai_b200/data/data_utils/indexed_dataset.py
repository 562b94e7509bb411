"""Token datasets stored as ``<prefix>.bin`` (raw token ids) + ``<prefix>.idx`` (index), bit-compatible
with the Megatron / fairseq formats the reference reads and writes
(reference libai/data/data_utils/indexed_dataset.py):

* legacy ``TNTIDX`` (:137-217): header ``TNTIDX\\0\\0 | u64 version=1 | u64 dtype code | u64 element size |
  u64 n_items | u64 n_sizes | u64 n_doc_idx`` followed by int64 ``dim_offsets[n_items+1]``,
  ``data_offsets[n_items+1]``, ``sizes[n_sizes]``, ``doc_idx[n_doc_idx]``; ``lazy`` reads items with
  seek/readinto, ``cached`` prefetches the requested items into one array (:220-269).
* ``MMIDIDX`` (:343-551): ``MMIDIDX\\0\\0 | u64 version=1 | u8 dtype code | u64 n_sizes | u64 n_docs`` then
  int32 ``sizes``, int64 byte ``pointers``, int64 ``doc_idx``; both files are memory mapped.

dtype codes: 1 u8, 2 i8, 3 i16, 4 i32, 5 i64, 6 f32, 7 f64, 8 u16.
"""
from __future__ import annotations

import logging
import os
import shutil
import struct
import time
from functools import lru_cache
from itertools import accumulate

import numpy as np
import torch

logger = logging.getLogger(__name__)

dtypes = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float32, 7: np.double, 8: np.uint16}
_LEGACY_MAGIC = b"TNTIDX\x00\x00"
_MMAP_MAGIC = b"MMIDIDX\x00\x00"


def code(dtype) -> int:
    for k, v in dtypes.items():
        if v == dtype:
            return k
    raise ValueError(dtype)


def index_file_path(prefix_path):
    return prefix_path + ".idx"


def data_file_path(prefix_path):
    return prefix_path + ".bin"


def best_fitting_dtype(vocab_size=None):
    return np.uint16 if vocab_size is not None and vocab_size < 65500 else np.int32


def get_available_dataset_impl():
    return ["lazy", "cached", "mmap"]


def infer_dataset_impl(path):
    if not IndexedDataset.exists(path):
        logger.info(f"Dataset does not exist: {path}")
        logger.info("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    with open(index_file_path(path), "rb") as f:
        head = f.read(8)
    if head == _LEGACY_MAGIC:
        return "cached"
    if head == _MMAP_MAGIC[:8]:
        return "mmap"
    return None


def make_builder(out_file, impl, vocab_size=None):
    if impl == "mmap":
        return MMapIndexedDatasetBuilder(out_file, dtype=best_fitting_dtype(vocab_size))
    return IndexedDatasetBuilder(out_file)


def make_dataset(path, impl, skip_warmup=False):
    if not IndexedDataset.exists(path):
        logger.info(f"Dataset does not exist: {path}")
        raise ValueError(f"Dataset does not exist: {path}")
    if impl == "infer":
        impl = infer_dataset_impl(path)
    if impl == "lazy":
        return IndexedDataset(path)
    if impl == "cached":
        return IndexedCachedDataset(path)
    if impl == "mmap" and MMapIndexedDataset.exists(path):
        return MMapIndexedDataset(path, skip_warmup)
    logger.info(f"Unknown dataset implementation: {impl}")
    return None


def dataset_exists(path, impl):
    return MMapIndexedDataset.exists(path) if impl == "mmap" else IndexedDataset.exists(path)


def read_longs(f, n):
    a = np.empty(n, dtype=np.int64)
    f.readinto(a)
    return a


def write_longs(f, a):
    f.write(np.array(a, dtype=np.int64))


def create_doc_idx(sizes):
    """Document boundaries when documents are separated by empty items."""
    return [0] + [i + 1 for i, s in enumerate(sizes) if s == 0]


# ---------------------------------------------------------------------------------------------------
# legacy format
# ---------------------------------------------------------------------------------------------------
class IndexedDataset(torch.utils.data.Dataset):
    """Lazy reader of the legacy format (one ``seek + readinto`` per item)."""

    _HDR_MAGIC = _LEGACY_MAGIC

    def __init__(self, path):
        super().__init__()
        self.path = path
        self.data_file = None
        self.read_index(path)

    def read_index(self, path):
        with open(index_file_path(path), "rb") as f:
            assert f.read(8) == self._HDR_MAGIC, (
                "Index file doesn't match expected format. Make sure that --dataset-impl is configured properly."
            )
            (version,) = struct.unpack("<Q", f.read(8))
            assert version == 1
            dcode, self.element_size = struct.unpack("<QQ", f.read(16))
            self.dtype = dtypes[dcode]
            self._len, self.s = struct.unpack("<QQ", f.read(16))
            (self.doc_count,) = struct.unpack("<Q", f.read(8))
            self.dim_offsets = read_longs(f, self._len + 1)
            self.data_offsets = read_longs(f, self._len + 1)
            self.sizes = read_longs(f, self.s)
            self.doc_idx = read_longs(f, self.doc_count)

    def read_data(self, path):
        self.data_file = open(data_file_path(path), "rb", buffering=0)

    def check_index(self, i):
        if i < 0 or i >= self._len:
            raise IndexError("index out of range")

    def __del__(self):
        if getattr(self, "data_file", None):
            self.data_file.close()

    def _read(self, start, stop):
        shape = self.sizes[self.dim_offsets[start] : self.dim_offsets[stop]]
        n = int(self.data_offsets[stop] - self.data_offsets[start])
        a = np.empty(n, dtype=self.dtype)
        self.data_file.seek(int(self.data_offsets[start]) * self.element_size)
        self.data_file.readinto(a)
        return a, shape

    def __getitem__(self, idx):
        if not self.data_file:
            self.read_data(self.path)
        if isinstance(idx, (int, np.integer)):
            self.check_index(idx)
            a, shape = self._read(idx, idx + 1)
            return a.reshape(tuple(int(x) for x in shape))
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("Slices into indexed_dataset must be contiguous")
            a, sizes = self._read(start, stop)
            return np.split(a, list(accumulate(int(s) for s in sizes))[:-1])
        raise TypeError(type(idx))

    def __len__(self):
        return self._len

    def num_tokens(self, index):
        return self.sizes[index]

    def size(self, index):
        return self.sizes[index]

    def get_doc_idx(self):
        return self.doc_idx

    def set_doc_idx(self, doc_idx_):
        self.doc_idx = doc_idx_

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))

    @property
    def supports_prefetch(self):
        return False


class IndexedCachedDataset(IndexedDataset):
    """Legacy format with an in-memory cache filled by ``prefetch(indices)``."""

    def __init__(self, path):
        super().__init__(path)
        self.cache = None
        self.cache_index = {}

    @property
    def supports_prefetch(self):
        return True

    def prefetch(self, indices):
        if all(i in self.cache_index for i in indices):
            return
        if not self.data_file:
            self.read_data(self.path)
        wanted = sorted(set(int(i) for i in indices))
        total = sum(int(self.data_offsets[i + 1] - self.data_offsets[i]) for i in wanted)
        self.cache = np.empty(total, dtype=self.dtype)
        self.cache_index.clear()
        cursor = 0
        for i in wanted:
            n = int(self.data_offsets[i + 1] - self.data_offsets[i])
            self.cache_index[i] = cursor
            self.data_file.seek(int(self.data_offsets[i]) * self.element_size)
            self.data_file.readinto(self.cache[cursor : cursor + n])
            cursor += n
        self.data_file.close()
        self.data_file = None

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            self.check_index(idx)
            shape = tuple(int(x) for x in self.sizes[self.dim_offsets[idx] : self.dim_offsets[idx + 1]])
            start = self.cache_index.get(int(idx))
            if start is None:          # not prefetched: read through from disk like the lazy dataset
                return super().__getitem__(idx)
            return self.cache[start : start + int(np.prod(shape))].reshape(shape).copy()
        if isinstance(idx, slice):
            return [self[i] for i in range(*idx.indices(len(self)))]
        raise TypeError(type(idx))


class IndexedDatasetBuilder:
    element_sizes = {np.uint8: 1, np.int8: 1, np.int16: 2, np.int32: 4, np.int64: 8, np.float32: 4, np.double: 8, np.uint16: 2}

    def __init__(self, out_file, dtype=np.int32):
        self.out_file = open(out_file, "wb")
        self.dtype = dtype
        self.data_offsets = [0]
        self.dim_offsets = [0]
        self.sizes = []
        self.element_size = self.element_sizes[self.dtype]
        self.doc_idx = [0]

    def add_item(self, tensor):
        arr = np.asarray(tensor.numpy() if hasattr(tensor, "numpy") else tensor, dtype=self.dtype)
        nbytes = self.out_file.write(arr.tobytes(order="C"))
        self.data_offsets.append(self.data_offsets[-1] + nbytes // self.element_size)
        self.sizes.extend(int(s) for s in arr.shape)
        self.dim_offsets.append(self.dim_offsets[-1] + arr.ndim)

    def end_document(self):
        self.doc_idx.append(len(self.sizes))

    def merge_file_(self, another_file):
        other = IndexedDataset(another_file)
        assert other.dtype == self.dtype
        base = self.data_offsets[-1]
        self.data_offsets.extend(base + int(o) for o in other.data_offsets[1:])
        self.sizes.extend(int(s) for s in other.sizes)
        base = self.dim_offsets[-1]
        self.dim_offsets.extend(base + int(o) for o in other.dim_offsets[1:])
        with open(data_file_path(another_file), "rb") as f:
            shutil.copyfileobj(f, self.out_file)

    def finalize(self, index_file):
        self.out_file.close()
        with open(index_file, "wb") as index:
            index.write(_LEGACY_MAGIC)
            index.write(struct.pack("<Q", 1))
            index.write(struct.pack("<QQ", code(self.dtype), self.element_size))
            index.write(struct.pack("<QQ", len(self.data_offsets) - 1, len(self.sizes)))
            index.write(struct.pack("<Q", len(self.doc_idx)))
            write_longs(index, self.dim_offsets)
            write_longs(index, self.data_offsets)
            write_longs(index, self.sizes)
            write_longs(index, self.doc_idx)


# ---------------------------------------------------------------------------------------------------
# mmap format
# ---------------------------------------------------------------------------------------------------
def _warmup_mmap_file(path):
    with open(path, "rb") as stream:
        while stream.read(100 * 1024 * 1024):
            pass


class MMapIndexedDataset(torch.utils.data.Dataset):
    class Index:
        _HDR_MAGIC = _MMAP_MAGIC

        @classmethod
        def writer(cls, path, dtype):
            class _Writer:
                def __enter__(self):
                    self._file = open(path, "wb")
                    self._file.write(cls._HDR_MAGIC)
                    self._file.write(struct.pack("<Q", 1))
                    self._file.write(struct.pack("<B", code(dtype)))
                    return self

                @staticmethod
                def _get_pointers(sizes):
                    itemsize = dtype().itemsize
                    ptrs = np.zeros(len(sizes), dtype=np.int64)
                    if len(sizes) > 1:
                        np.cumsum(np.asarray(sizes[:-1], dtype=np.int64) * itemsize, out=ptrs[1:])
                    return ptrs

                def write(self, sizes, doc_idx):
                    self._file.write(struct.pack("<Q", len(sizes)))
                    self._file.write(struct.pack("<Q", len(doc_idx)))
                    self._file.write(np.array(sizes, dtype=np.int32).tobytes(order="C"))
                    self._file.write(self._get_pointers(sizes).tobytes(order="C"))
                    self._file.write(np.array(doc_idx, dtype=np.int64).tobytes(order="C"))

                def __exit__(self, exc_type, exc_val, exc_tb):
                    self._file.close()

            return _Writer()

        def __init__(self, path, skip_warmup=False):
            with open(path, "rb") as stream:
                assert stream.read(9) == self._HDR_MAGIC, (
                    "Index file doesn't match expected format. Make sure that --dataset-impl is configured properly."
                )
                (version,) = struct.unpack("<Q", stream.read(8))
                assert version == 1
                (dcode,) = struct.unpack("<B", stream.read(1))
                self._dtype = dtypes[dcode]
                self._dtype_size = self._dtype().itemsize
                (self._len,) = struct.unpack("<Q", stream.read(8))
                (self._doc_count,) = struct.unpack("<Q", stream.read(8))
                offset = stream.tell()
            if not skip_warmup:
                logger.info("warming up index mmap file...")
                _warmup_mmap_file(path)
            self._bin_buffer_mmap = np.memmap(path, mode="r", order="C")
            self._bin_buffer = memoryview(self._bin_buffer_mmap)
            logger.info("reading sizes...")
            self._sizes = np.frombuffer(self._bin_buffer, dtype=np.int32, count=self._len, offset=offset)
            logger.info("reading pointers...")
            self._pointers = np.frombuffer(self._bin_buffer, dtype=np.int64, count=self._len, offset=offset + self._sizes.nbytes)
            logger.info("reading document index...")
            self._doc_idx = np.frombuffer(
                self._bin_buffer, dtype=np.int64, count=self._doc_count,
                offset=offset + self._sizes.nbytes + self._pointers.nbytes,
            )

        def __del__(self):
            if hasattr(self, "_bin_buffer_mmap"):
                self._bin_buffer_mmap._mmap.close()
                del self._bin_buffer_mmap

        @property
        def dtype(self):
            return self._dtype

        @property
        def sizes(self):
            return self._sizes

        @property
        def doc_idx(self):
            return self._doc_idx

        @lru_cache(maxsize=8)
        def __getitem__(self, i):
            return self._pointers[i], self._sizes[i]

        def __len__(self):
            return self._len

    def __init__(self, path, skip_warmup=False):
        super().__init__()
        self._path = None
        self._index = None
        self._bin_buffer = None
        self._do_init(path, skip_warmup)

    def __getstate__(self):
        return self._path

    def __setstate__(self, state):
        self._do_init(state, skip_warmup=True)

    def _do_init(self, path, skip_warmup):
        self._path = path
        self._index = self.Index(index_file_path(path), skip_warmup)
        if not skip_warmup:
            logger.info("warming up data mmap file...")
            _warmup_mmap_file(data_file_path(path))
        logger.info("creating numpy buffer of mmap...")
        self._bin_buffer_mmap = np.memmap(data_file_path(path), mode="r", order="C")
        logger.info("creating memory view of numpy buffer...")
        self._bin_buffer = memoryview(self._bin_buffer_mmap)

    def __del__(self):
        if getattr(self, "_bin_buffer_mmap", None) is not None:
            self._bin_buffer_mmap._mmap.close()
            del self._bin_buffer_mmap
        self._index = None

    def __len__(self):
        return len(self._index)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            ptr, size = self._index[int(idx)]
            return np.frombuffer(self._bin_buffer, dtype=self._index.dtype, count=size, offset=ptr)
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("Slices into indexed_dataset must be contiguous")
            ptr = self._index._pointers[start]
            sizes = self._index._sizes[idx]
            total = int(sizes.sum())
            flat = np.frombuffer(self._bin_buffer, dtype=self._index.dtype, count=total, offset=ptr)
            return np.split(flat, list(accumulate(int(s) for s in sizes))[:-1])
        raise TypeError(type(idx))

    def get(self, idx, offset=0, length=None):
        """A slice ``[offset, offset + length)`` of item ``idx`` without touching the rest."""
        ptr, size = self._index[int(idx)]
        if length is None:
            length = size - offset
        ptr += offset * np.dtype(self._index.dtype).itemsize
        return np.frombuffer(self._bin_buffer, dtype=self._index.dtype, count=length, offset=ptr)

    @property
    def sizes(self):
        return self._index.sizes

    @property
    def doc_idx(self):
        return self._index.doc_idx

    def get_doc_idx(self):
        return self._index._doc_idx

    def set_doc_idx(self, doc_idx_):
        self._index._doc_idx = doc_idx_

    @property
    def supports_prefetch(self):
        return False

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))


class MMapIndexedDatasetBuilder:
    def __init__(self, out_file, dtype=np.int64):
        self._data_file = open(out_file, "wb")
        self._dtype = dtype
        self._sizes = []
        self._doc_idx = [0]

    def add_item(self, tensor):
        arr = np.asarray(tensor.numpy() if hasattr(tensor, "numpy") else tensor, dtype=self._dtype)
        self._data_file.write(arr.tobytes(order="C"))
        self._sizes.append(arr.size)

    def end_document(self):
        self._doc_idx.append(len(self._sizes))

    def merge_file_(self, another_file):
        index = MMapIndexedDataset.Index(index_file_path(another_file), skip_warmup=True)
        assert index.dtype == self._dtype
        self._sizes.extend(int(s) for s in index.sizes)
        with open(data_file_path(another_file), "rb") as f:
            shutil.copyfileobj(f, self._data_file)

    def finalize(self, index_file):
        self._data_file.close()
        with MMapIndexedDataset.Index.writer(index_file, self._dtype) as index:
            index.write(self._sizes, self._doc_idx)


def get_indexed_dataset(data_prefix, data_impl, skip_warmup):
    """Open ``data_prefix`` with implementation ``data_impl`` ("mmap" | "lazy" | "cached" | "infer")."""
    logger.info("building dataset index ...")
    t0 = time.time()
    ds = make_dataset(data_prefix, data_impl, skip_warmup)
    assert ds is not None and ds.sizes.shape[0] == ds.doc_idx[-1]
    logger.info("Finished creating indexed dataset in {:4f} seconds".format(time.time() - t0))
    logger.info("indexed dataset stats:")
    logger.info("number of documents: {}".format(ds.doc_idx.shape[0] - 1))
    logger.info("number of sentences: {}".format(ds.sizes.shape[0]))
    return ds
