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
    restores the caller's order (e.g. for the tokens of `sentence_index`)."""
    dev = captions.device
    L = torch.as_tensor(lengths, dtype=torch.int64, device=dev)
    if L.dim() != 1 or L.shape[0] != captions.shape[0] or features.shape[0] != captions.shape[0]:
        raise ValueError("sort_batch: features, captions and lengths must agree on the batch size")
    if L.numel() == 0:
        raise RuntimeError("empty batch")
    Ls, perm = torch.sort(L, descending=True, stable=True)
    lens = [int(x) for x in Ls.tolist()]                       # the only device -> host traffic: B integers
    if lens[-1] <= 0:
        raise RuntimeError("Length of all samples has to be greater than 0")
    if lens[0] > captions.shape[1]:
        raise ValueError("sort_batch: a length exceeds the padded caption width")
    return (features.index_select(0, perm), captions.index_select(0, perm)[:, :lens[0]].contiguous(), lens, perm)
