"""Collate / packing contract of the reference's data loader (utils.py:61-77 `create_batch`): the batch handed to
the decoder is sorted by caption length, longest first (Python's stable sort, `reverse=True`), captions are
zero-padded to the longest one, and the lengths travel as a Python list -- `pack_padded_sequence(...,
enforce_sorted=True)` (rnn.py:31) and therefore this package's decoders depend on exactly that order.

`create_batch` is the drop-in for `utils.create_batch` (same input list, same outputs).  `sort_batch` is the same
contract for tensors that already live on the GPU (features from a device-side encoder, a padded caption matrix in
arbitrary order): one stable sort of the B lengths and row gathers on the device, so callers no longer have to
pre-sort.  Only the B lengths cross to the host (the kernels take `batch_sizes` by value).  The packed targets
themselves are built on the device by the decoders (`st_pack_targets`, main.py:145).
"""
import torch


def create_batch(data):
    """utils.py:61-77.  data: list of (image_path, image (C,H,W), caption (len,) int64) -> (paths, images (B,C,H,W),
    target_captions (B, max_len) int64 zero-padded, caption_len list), sorted by caption length, longest first,
    ties in their original order.  (Unlike the reference it does not sort the caller's list in place.)"""
    order = sorted(range(len(data)), key=lambda i: len(data[i][2]), reverse=True)
    image_paths = tuple(data[i][0] for i in order)
    images = torch.stack([data[i][1] for i in order], 0)
    captions = [data[i][2] for i in order]
    caption_len = [len(c) for c in captions]
    target = torch.zeros(len(captions), max(caption_len), dtype=torch.long)
    for idx, c in enumerate(captions):
        target[idx, :caption_len[idx]] = c[:caption_len[idx]]
    return image_paths, images, target, caption_len


def sort_batch(features, captions, lengths):
    """features (B, ...), captions (B, T) int64 zero-padded, lengths (B,) tensor or list, in ANY order ->
    (features, captions[:, :max_len], lengths list, perm) in the decoder's order (length-descending, stable).
    `perm` (B,) int64 on the tensors' device maps sorted row i to original row perm[i]: `out[perm] = sorted_out`
    restores the caller's order (e.g. for the tokens of `sentence_index`).
    CUDA tensors: the library's collate kernels (st_collate_sort: stable rank by counting, st_gather_rows_bytes,
    st_collate_captions); CPU tensors (host-side use, tests): torch.sort + index_select."""
    dev = captions.device
    L = torch.as_tensor(lengths, dtype=torch.int64, device=dev)
    if L.dim() != 1 or L.shape[0] != captions.shape[0] or features.shape[0] != captions.shape[0]:
        raise ValueError("sort_batch: features, captions and lengths must agree on the batch size")
    if L.numel() == 0:
        raise RuntimeError("empty batch")
    if captions.is_cuda:
        return _sort_batch_cuda(features, captions, L)
    Ls, perm = torch.sort(L, descending=True, stable=True)
    lens = [int(x) for x in Ls.tolist()]                       # the only device -> host traffic: B integers
    _check_lengths(lens, captions.shape[1])
    return (features.index_select(0, perm), captions.index_select(0, perm)[:, :lens[0]].contiguous(), lens, perm)


def _check_lengths(lens, width):
    if lens[-1] <= 0:
        raise RuntimeError("Length of all samples has to be greater than 0")
    if lens[0] > width:
        raise ValueError("sort_batch: a length exceeds the padded caption width")


def _sort_batch_cuda(features, captions, L):
    from . import _lib
    from ._lib import check, ptr, stream_ptr
    lib = _lib.load()
    if not features.is_cuda:
        raise RuntimeError("sort_batch: features and captions must live on the same device")
    B, T = captions.shape
    if B > 65535:
        raise ValueError("sort_batch: at most 65535 samples per batch")
    dev = captions.device
    captions = captions.contiguous()
    features = features.contiguous()
    L = L.contiguous()
    perm = torch.empty(B, dtype=torch.int64, device=dev)
    Ls = torch.empty(B, dtype=torch.int64, device=dev)
    check(lib.st_collate_sort(ptr(L), B, 0, ptr(perm), ptr(Ls), None, stream_ptr()), "st_collate_sort")
    lens = [int(x) for x in Ls.tolist()]                       # the only device -> host traffic: B integers
    _check_lengths(lens, T)
    cap = torch.empty(B, lens[0], dtype=torch.int64, device=dev)
    check(lib.st_collate_captions(ptr(cap), ptr(captions), ptr(perm), ptr(Ls), B, T, lens[0], stream_ptr()), "st_collate_captions")
    feat = torch.empty_like(features)
    row_bytes = features[0].numel() * features.element_size()
    check(lib.st_gather_rows_bytes(ptr(feat), ptr(features), ptr(perm), B, row_bytes, stream_ptr()), "st_gather_rows_bytes")
    return feat, cap, lens, perm


def device_batch_sizes(lengths_dev, T):
    """batch_sizes (T,) int32 on the device for (unsorted) device lengths: #{i : len_i > t} (rnn.py:31's
    pack_padded_sequence derives the same numbers from the sorted lengths)."""
    from . import _lib
    from ._lib import check, ptr, stream_ptr
    lib = _lib.load()
    L = lengths_dev.to(torch.int64).contiguous()
    B = L.shape[0]
    perm = torch.empty(B, dtype=torch.int64, device=L.device)
    Ls = torch.empty(B, dtype=torch.int64, device=L.device)
    bs = torch.empty(T, dtype=torch.int32, device=L.device)
    check(lib.st_collate_sort(ptr(L), B, int(T), ptr(perm), ptr(Ls), ptr(bs), stream_ptr()), "st_collate_sort")
    return bs
