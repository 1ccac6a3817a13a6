"""Streaming input for the GPU scanner.

`Slicer` mirrors the geometry of the reference's input iterator (/root/reference/src/input.rs:33-168): the inputs are
read one after the other in pieces of at most INPUT_BUF_LEN = 4096 bytes, every file starts a new 4096-byte grid,
an empty piece marks the switch to the next file, the file label is 1-based (None for stdin) and -- a quirk the
scanner's results depend on -- the "this is the very last piece" flag never reaches the consumer as true
(input.rs:130-137 returns None in exactly that case).

`scan_files` is the driver of main.rs:143-167 for the GPU: instead of one `FindingCollection::from` per 4096-byte
slice it hands the library large pieces (`sx_scan_stream` folds over the slices inside the call) and keeps the GPU
fed: two pinned staging buffers and two device buffers per GPU, piece k+1 is read from disk and copied to the device
on a copy stream while piece k is being scanned.  torch is used for the pinned / device memory and the streams only.
"""
from __future__ import annotations

import sys
from typing import BinaryIO, Iterator, List, Optional, Sequence, Tuple

from .scanner import FindingCollection, ScannerState

INPUT_BUF_LEN = 4096  # input.rs:22


class Slicer:
    """input.rs:33-168 as a Python iterator: yields (slice, input_file_id, is_last) with the reference's geometry."""

    def __init__(self, inputs: Sequence[str], stdin: Optional[BinaryIO] = None):
        self.from_stdin = len(inputs) == 0 or (len(inputs) == 1 and inputs[0] == "-")  # input.rs:64-66
        self.inputs = list(inputs)
        self.stdin = stdin if stdin is not None else sys.stdin.buffer

    @staticmethod
    def _open(path: str) -> Optional[BinaryIO]:
        try:
            return open(path, "rb")
        except OSError as e:  # input.rs:78-84: reported, then treated as an empty input
            print(f"Error: can not read file `{path}`: {e}", file=sys.stderr)
            return None

    def __iter__(self) -> Iterator[Tuple[bytes, Optional[int], bool]]:
        if self.from_stdin:
            while True:
                b = self.stdin.read(INPUT_BUF_LEN)
                if not b:
                    return  # input.rs:130-137: the consumer never sees is_last == true
                yield b, None, False
        idx, reader = 1, self._open(self.inputs[0])
        while True:
            b = reader.read(INPUT_BUF_LEN) if reader else b""
            if not b:
                if idx == len(self.inputs):
                    return
                idx += 1  # the empty piece already carries the next file's label (input.rs:138-161)
                reader = self._open(self.inputs[idx - 1])
            yield b, idx, False


def scan_files(states: Sequence[ScannerState], inputs: Sequence[str], chunk_bytes: int = 256 << 20,
               stdin: Optional[BinaryIO] = None) -> Iterator[Tuple[Optional[int], List[FindingCollection]]]:
    """Scan the concatenated inputs with every state (one per mission; they may sit on different GPUs).

    Yields (input_file_id, [one FindingCollection per state]) per piece of at most chunk_bytes (a multiple of 4096, so
    the slice grid inside the library is the reference's); merging a piece's collections with `merge` gives the order
    of main.rs:118-136.  Equivalent to feeding `Slicer`'s slices one by one to `ScannerState.scan` (tests compare)."""
    import torch

    assert chunk_bytes % INPUT_BUF_LEN == 0 and chunk_bytes > 0
    from_stdin = len(inputs) == 0 or (len(inputs) == 1 and inputs[0] == "-")
    devices = sorted({s.device for s in states})
    host = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    hview = [memoryview(h.numpy()) for h in host]
    dev = {d: [torch.empty(chunk_bytes, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(2)] for d in devices}
    copy_stream = {d: torch.cuda.Stream(device=d) for d in devices}
    copied = {d: [None, None] for d in devices}   # event: piece in dev[d][k] is complete
    scanned = {d: [None, None] for d in devices}  # event: dev[d][k] may be overwritten
    host_free = [None, None]                      # events: host[k] has been copied to every device

    def pieces():
        """(file id, reader) pairs in input order; an unreadable file is an empty input."""
        if from_stdin:
            yield None, (stdin if stdin is not None else sys.stdin.buffer)
            return
        for i, p in enumerate(inputs, start=1):
            yield i, Slicer._open(p)

    def fill(reader, k) -> int:
        """Read up to chunk_bytes into host[k] (short reads are retried: only EOF ends a piece early)."""
        if host_free[k] is not None:
            for ev in host_free[k]:
                ev.synchronize()
            host_free[k] = None
        n = 0
        while reader is not None and n < chunk_bytes:
            got = reader.readinto(hview[k][n:])
            if not got:
                break
            n += got
        return n

    def upload(k, n):
        evs = []
        for d in devices:
            with torch.cuda.device(d), torch.cuda.stream(copy_stream[d]):
                if scanned[d][k] is not None:
                    copy_stream[d].wait_event(scanned[d][k])
                dev[d][k][:n].copy_(host[k][:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream[d])
                copied[d][k] = ev
                evs.append(ev)
        host_free[k] = evs

    for fid, reader in pieces():
        k = 0
        n = fill(reader, k)
        if n:
            upload(k, n)
        while n:
            # next piece of the same file: disk read + H2D overlap the scan of the current one
            k2 = k ^ 1
            n2 = fill(reader, k2) if n == chunk_bytes else 0
            if n2:
                upload(k2, n2)
            out = []
            for s in states:
                d = s.device
                with torch.cuda.device(d):
                    torch.cuda.current_stream(d).wait_event(copied[d][k])
                    out.append(s.scan_stream(None, False, INPUT_BUF_LEN, fid, device_ptr=dev[d][k].data_ptr(), length=n,
                                             cuda_stream=torch.cuda.current_stream(d).cuda_stream))
            for d in devices:
                with torch.cuda.device(d):
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(d))
                    scanned[d][k] = ev
            yield fid, out
            k, n = k2, n2
        if reader is not None and not from_stdin:
            reader.close()
