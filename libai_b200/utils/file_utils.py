"""Cache-aware fetching of datasets / vocab files.

Spec: reference libai/utils/file_utils.py — ``url_to_filename`` (:39), ``cached_path`` (:90),
``get_from_cache`` (:189), ``get_md5`` (:267), ``get_data_from_cache(url, cache_dir, md5)``
(:281).  S3 access requires boto3 (not in the image) and raises a clear error.
"""
import hashlib
import json
import logging
import os
import shutil
import tempfile
from pathlib import Path
from urllib.parse import urlparse

from . import distributed as dutil
from .file_io import get_cache_dir

logger = logging.getLogger(__name__)
DEFAULT_CACHE_DIR = os.getenv("LIBAI_DATA_CACHE", os.path.join(get_cache_dir(), "data"))


def url_to_filename(url: str, etag: str = None) -> str:
    name = hashlib.sha256(url.encode("utf-8")).hexdigest()
    if etag:
        name += "." + hashlib.sha256(etag.encode("utf-8")).hexdigest()
    return name


def filename_to_url(filename: str, cache_dir=None):
    cache_dir = str(cache_dir or DEFAULT_CACHE_DIR)
    cache_path = os.path.join(cache_dir, filename)
    if not os.path.exists(cache_path):
        raise FileNotFoundError(f"file {cache_path} not found")
    meta_path = cache_path + ".json"
    if not os.path.exists(meta_path):
        raise FileNotFoundError(f"file {meta_path} not found")
    with open(meta_path) as f:
        meta = json.load(f)
    return meta["url"], meta["etag"]


def cached_path(url_or_filename, cache_dir=None) -> str:
    """Local path for a URL (downloaded into the cache) or an existing file."""
    url_or_filename = str(url_or_filename)
    cache_dir = str(cache_dir or DEFAULT_CACHE_DIR)
    scheme = urlparse(url_or_filename).scheme
    if scheme in ("http", "https", "s3"):
        return get_from_cache(url_or_filename, cache_dir)
    if os.path.exists(url_or_filename):
        return url_or_filename
    if scheme == "":
        raise FileNotFoundError(f"file {url_or_filename} not found")
    raise ValueError(f"unable to parse {url_or_filename} as a URL or as a local path")


def split_s3_path(url: str):
    parsed = urlparse(url)
    if not parsed.netloc or not parsed.path:
        raise ValueError(f"bad s3 path {url}")
    return parsed.netloc, parsed.path.lstrip("/")


def _boto3():
    try:
        import boto3  # noqa

        return boto3
    except ImportError as e:
        raise ImportError("s3:// paths need boto3, which is not installed in this image") from e


def s3_request(func):
    """Decorator for s3 calls: a 404 from the service becomes ``EnvironmentError("file … not found")``
    (reference libai/utils/file_utils.py:141-157)."""
    import functools

    @functools.wraps(func)
    def wrapper(url, *args, **kwargs):
        try:
            return func(url, *args, **kwargs)
        except Exception as exc:  # botocore.exceptions.ClientError when boto3 is present
            code = getattr(exc, "response", {}).get("Error", {}).get("Code") if hasattr(exc, "response") else None
            if code is not None and str(code) == "404":
                raise EnvironmentError(f"file {url} not found") from exc
            raise

    return wrapper


@s3_request
def s3_etag(url: str):
    bucket, key = split_s3_path(url)
    return _boto3().resource("s3").Object(bucket, key).e_tag


@s3_request
def s3_get(url: str, temp_file) -> None:
    bucket, key = split_s3_path(url)
    _boto3().resource("s3").Bucket(bucket).download_fileobj(key, temp_file)


def http_get(url: str, temp_file) -> None:
    import requests

    req = requests.get(url, stream=True)
    total = int(req.headers.get("Content-Length") or 0)
    try:
        from tqdm import tqdm

        bar = tqdm(unit="B", total=total or None)
    except ImportError:
        bar = None
    for chunk in req.iter_content(chunk_size=1 << 16):
        if chunk:
            temp_file.write(chunk)
            if bar:
                bar.update(len(chunk))
    if bar:
        bar.close()


def get_from_cache(url: str, cache_dir=None) -> str:
    import requests

    cache_dir = str(cache_dir or DEFAULT_CACHE_DIR)
    os.makedirs(cache_dir, exist_ok=True)
    if url.startswith("s3://"):
        etag = s3_etag(url)
    else:
        try:
            resp = requests.head(url, allow_redirects=True)
            etag = resp.headers.get("ETag") if resp.status_code == 200 else None
        except EnvironmentError:
            etag = None
    cache_path = os.path.join(cache_dir, url_to_filename(url, etag))
    if not os.path.exists(cache_path) and etag is None:
        # offline: fall back to any previously cached variant of this url
        stem = url_to_filename(url)
        matches = [f for f in os.listdir(cache_dir) if f.startswith(stem) and not f.endswith(".json")]
        if matches:
            cache_path = os.path.join(cache_dir, matches[-1])
    if not os.path.exists(cache_path):
        with tempfile.NamedTemporaryFile() as tmp:
            logger.info(f"{url} not found in cache, downloading to {tmp.name}")
            (s3_get if url.startswith("s3://") else http_get)(url, tmp)
            tmp.flush()
            tmp.seek(0)
            with open(cache_path, "wb") as out:
                shutil.copyfileobj(tmp, out)
            with open(cache_path + ".json", "w") as mf:
                json.dump({"url": url, "etag": etag}, mf)
    return cache_path


def get_md5(fname: str) -> str:
    h = hashlib.md5()
    with open(fname, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def download_file(out_path: str, url: str) -> None:
    logger.info(f"downloading from {url} to {out_path}")
    with open(out_path, "wb") as f:
        http_get(url, f)


def get_data_from_cache(url: str, cache_dir=None, md5: str = None) -> str:
    """Return ``cache_dir/<basename(url)>``, downloading on local-rank 0 when missing or when the
    md5 differs; other ranks wait on the barrier."""
    cache_dir = Path(cache_dir or DEFAULT_CACHE_DIR)
    cache_dir.mkdir(parents=True, exist_ok=True)
    target = cache_dir / url.split("/")[-1]
    if dutil.get_local_rank() == 0:
        if target.exists() and md5 is not None and get_md5(str(target)) != md5:
            os.unlink(target)
        if not target.exists():
            download_file(str(target), url)
    dutil.synchronize()
    assert target.exists(), f"{target} is missing and could not be downloaded (no network?)"
    if md5 is not None:
        got = get_md5(str(target))
        assert got == md5, f"{target} md5 mismatch: {got} != {md5}"
    return str(target)
