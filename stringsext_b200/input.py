"""Streaming input for the GPU scanner.

`Slicer` mirrors the geometry of the reference's input iterator (/root/reference/src/input.rs:33-168): the inputs are
read one after the other in pieces of at most INPUT_BUF_LEN = 4096 bytes, every file starts a new 4096-byte grid,
an empty piece marks the switch to the next file, the file label is 1-based (None for stdin) and -- a quirk the
scanner's results depend on -- the "this is the very last piece" flag never reaches the consumer as true
(input.rs:130-137 returns None in exactly that case).

`scan_inputs` / `scan_files` drive main.rs:143-167 for the GPU through the library's C streaming driver
(`sx_scan_file` / `sx_scan_reader`, include/stringsext_b200.h): instead of one `FindingCollection::from` per 4096-byte
slice the library gets large pieces (`sx_scan_stream` folds over the slices inside the call) and keeps the GPUs fed:
two pinned staging buffers and two device buffers per GPU, piece k+1 is read and copied to the devices while piece k
is being scanned, and the missions' scans run side by side (one worker thread per ScannerState).  No torch involved.
"""
from __future__ import annotations

import sys
from typing import BinaryIO, Callable, Iterator, List, Optional, Sequence, Tuple

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
                # input.rs:118 is one Read::read() per slice: a short read from a pipe is a slice of its own
                b = self.stdin.read1(INPUT_BUF_LEN) if hasattr(self.stdin, "read1") else self.stdin.read(INPUT_BUF_LEN)
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


def scan_inputs(states: Sequence[ScannerState], inputs: Sequence[str], on_batch: Callable[[Optional[int], List[FindingCollection]], None],
                chunk_bytes: int = 256 << 20, stdin: Optional[BinaryIO] = None) -> None:
    """Stream the concatenated inputs through every state (one per mission; they may sit on different GPUs) with the
    library's streaming driver `sx_scan_file` / `sx_scan_reader` (C: pinned double buffering, one upload per device,
    the states' scans side by side on their worker threads, read + upload of piece k + 1 beside the scan of piece k).

    `on_batch(input_file_id, [one FindingCollection per state])` is called per piece of at most chunk_bytes (a multiple of
    4096, so the slice grid inside the library is the reference's); merging a piece's collections with `merge` gives the
    order of main.rs:118-136.  Equivalent to feeding `Slicer`'s slices one by one to `ScannerState.scan` (tests compare).
    Inputs read from `stdin` (or a Python file object) go through the reader-callback form; note that the library
    normalises the grid to full 4096-byte slices, while the reference's `Read::read` (input.rs:118) turns every short
    read of a pipe into its own slice -- only regular files are reproducible in the reference."""
    import ctypes as C

    from .scanner import BATCH_FN, READ_FN, load_library

    assert chunk_bytes % INPUT_BUF_LEN == 0 and chunk_bytes > 0
    L = load_library()
    harr = (C.c_void_p * len(states))(*[s._h for s in states])
    err: list = []

    def batch(_user, fid, fcs, n):
        try:
            file_id = None if fid < 0 else fid
            on_batch(file_id, [states[i]._collect(fcs[i], fid) for i in range(n)])
            return 0
        except BaseException as e:  # noqa: BLE001 -- must not unwind through C
            err.append(e)
            return 1

    cb = BATCH_FN(batch)
    from_stdin = len(inputs) == 0 or (len(inputs) == 1 and inputs[0] == "-")
    if from_stdin:
        rd = stdin if stdin is not None else sys.stdin.buffer

        def read(_user, dst, cap):
            b = rd.read(cap)
            if b:
                C.memmove(dst, b, len(b))
            return len(b)

        rc = L.sx_scan_reader(harr, len(states), -1, READ_FN(read), None, chunk_bytes, cb, None)
        if err:
            raise err[0]
        if rc < 0:
            from .scanner import _raise_last

            _raise_last()
        return
    for i, p in enumerate(inputs, start=1):
        rc = L.sx_scan_file(harr, len(states), i, p.encode(), chunk_bytes, cb, None)
        if err:
            raise err[0]
        if rc < 0:
            from .scanner import _raise_last

            _raise_last()


def scan_files(states: Sequence[ScannerState], inputs: Sequence[str], chunk_bytes: int = 256 << 20,
               stdin: Optional[BinaryIO] = None) -> Iterator[Tuple[Optional[int], List[FindingCollection]]]:
    """Generator form of `scan_inputs`: yields (input_file_id, [one FindingCollection per state]) per piece.  The C driver
    calls back and cannot be suspended, so the pieces are collected first; use `scan_inputs` with a callback for inputs
    whose findings do not fit in memory."""
    out: list = []
    scan_inputs(states, inputs, lambda fid, fcs: out.append((fid, fcs)), chunk_bytes, stdin)
    yield from out
