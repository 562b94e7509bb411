"""URL download with progress (spec: reference libai/utils/download.py:29-94)."""
import logging
import os
import shutil
from typing import Optional
from urllib import request


def download(url: str, dir: str, *, filename: Optional[str] = None, progress: bool = True) -> str:
    """Download ``url`` into ``dir`` (created if needed) and return the local path.
    An existing target is reused."""
    os.makedirs(dir, exist_ok=True)
    if filename is None:
        filename = url.split("/")[-1]
        assert len(filename), "Cannot obtain filename from url {}".format(url)
    fpath = os.path.join(dir, filename)
    logger = logging.getLogger(__name__)
    if os.path.isfile(fpath):
        logger.info(f"File {filename} exists! Skipping download.")
        return fpath
    tmp = fpath + ".tmp"
    try:
        logger.info(f"Downloading from {url} ...")
        bar = None
        if progress:
            try:
                import tqdm

                bar = tqdm.tqdm(unit="B", unit_scale=True, miniters=1, desc=filename, leave=True)
            except ImportError:
                bar = None

        def hook(blocks, bsize, total):
            if bar is not None:
                if total > 0:
                    bar.total = total
                bar.update(blocks * bsize - bar.n)

        tmp, _ = request.urlretrieve(url, filename=tmp, reporthook=hook)
        if bar is not None:
            bar.close()
        size = os.stat(tmp).st_size
        if size == 0:
            raise IOError(f"Downloaded an empty file from {url}!")
        shutil.move(tmp, fpath)
    finally:
        try:
            os.unlink(tmp)
        except (IOError, OSError):
            pass
    logger.info(f"Successfully downloaded {fpath}. {os.stat(fpath).st_size} bytes.")
    return fpath
