"""Cross-host weight transfer over N parallel TCP streams (for elastic REMOTE rollout instances that are not on
the trainer's NVSwitch domain).  Capability parity with the reference's TCPTransferEngine + sender/receiver
agents (rlboost/weight_transfer/transfer_engine.py:14-291: parallel streams, 16 MiB socket buffers, (offset, length)
header — extended here with an auth token, a transfer id and the total length so that a partly failed multi-stream transfer can
never be completed by streams of the NEXT one, and a stranger that can reach the ports cannot overwrite weights — ``os.sendfile`` from a memfd, ``recv_into`` a registered buffer, per-batch status;
sender_agent.py:33-99 memfd/shm staging; receiver_agent.py:22-32 shared buffer).  Same-box instances use the
peer-memory path in ``sync.py`` instead.
"""
from __future__ import annotations

import hashlib
import logging
import mmap
import os
import socket
import struct
import threading

import torch

HEADER = struct.Struct("<QQQQQ")       # (token digest, transfer id, total bytes of the transfer, offset, length)
SOCK_BUF = 16 << 20
log = logging.getLogger(__name__)


def token_digest(token: str | None) -> int:
    """64-bit digest of the shared secret (``POLYRL_WEIGHT_TOKEN`` unless given); both ends must agree."""
    token = os.environ.get("POLYRL_WEIGHT_TOKEN", "") if token is None else token
    return int.from_bytes(hashlib.sha256(token.encode()).digest()[:8], "little")


def flat_state_nbytes(named: list[tuple[str, torch.Tensor]]) -> int:
    return sum(t.numel() * t.element_size() for _, t in named)


class StagingBuffer:
    """memfd-backed host buffer: mmap for the D2H copies, file descriptor for zero-copy ``sendfile``."""

    def __init__(self, nbytes: int, name: str = "polyrl_weights"):
        self.nbytes = nbytes
        self.fd = os.memfd_create(name)
        os.ftruncate(self.fd, nbytes)
        self.mm = mmap.mmap(self.fd, nbytes)
        self.tensor = torch.frombuffer(self.mm, dtype=torch.uint8)
        if torch.cuda.is_available():
            try:
                torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)    # pinned => async D2H
            except Exception:
                pass

    def close(self):
        try:
            self.mm.close()
        finally:
            os.close(self.fd)


class TCPWeightSender:
    def __init__(self, num_streams: int = 4, token: str | None = None):
        self.num_streams = num_streams
        self.status: dict[int, str] = {}
        self._batch = 0
        self._token = token_digest(token)
        self._nonce = int.from_bytes(os.urandom(4), "little") << 32      # transfer ids are unique across sender restarts

    def stage(self, named: list[tuple[str, torch.Tensor]], staging: StagingBuffer) -> list[tuple[str, int, int]]:
        """Copy (full) tensors into the staging buffer at running offsets; returns the layout meta."""
        off, meta = 0, []
        for name, t in named:
            n = t.numel() * t.element_size()
            staging.tensor[off:off + n].copy_(t.detach().contiguous().view(torch.uint8).reshape(-1), non_blocking=True)
            meta.append((name, off, n))
            off += n
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return meta

    def _send_range(self, host: str, port: int, fd: int, lo: int, hi: int, zero_copy: bool, mm, tid: int = 0,
                    total: int = 0) -> None:
        s = socket.create_connection((host, port), timeout=60)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, SOCK_BUF)
        s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        s.sendall(HEADER.pack(self._token, tid, total, lo, hi - lo))
        sent = lo
        while sent < hi:
            if zero_copy:
                n = os.sendfile(s.fileno(), fd, sent, min(hi - sent, 64 << 20))
            else:
                n = s.send(memoryview(mm)[sent:min(hi, sent + (16 << 20))])
            if n == 0:
                raise ConnectionError("peer closed during weight transfer")
            sent += n
        ack = s.recv(2)
        s.close()
        if ack != b"OK":
            raise ConnectionError("receiver did not acknowledge")

    def send(self, staging: StagingBuffer, host: str, base_port: int, nbytes: int | None = None) -> int:
        """Split [0, nbytes) over the streams (receiver listens on base_port + i).  Returns a batch id."""
        nbytes = nbytes or staging.nbytes
        self._batch += 1
        bid = self._batch
        tid = self._nonce | (bid & 0xFFFFFFFF)
        self.status[bid] = "running"
        per = (nbytes + self.num_streams - 1) // self.num_streams
        per = (per + 4095) // 4096 * 4096
        zero_copy = os.environ.get("TCP_ZEROCOPY", "1") != "0"
        errs: list = []

        def work(i):
            lo, hi = i * per, min(nbytes, (i + 1) * per)
            if lo >= hi:
                return
            try:
                self._send_range(host, base_port + i, staging.fd, lo, hi, zero_copy, staging.mm, tid, nbytes)
            except Exception as e:
                errs.append(e)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(self.num_streams)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        self.status[bid] = "failed" if errs else "success"
        if errs:
            raise errs[0]
        return bid


class TCPWeightReceiver:
    """Listens on ``port + i`` for stream i, receives into a host buffer, then copies into the model tensors."""

    def __init__(self, weights, bind_host: str | None = None, port: int = 21000, num_streams: int = 4,
                 token: str | None = None):
        # bind to the registered NIC when the launcher names one (POLYRL_WEIGHT_BIND); every stream must present the token
        bind_host = bind_host or os.environ.get("POLYRL_WEIGHT_BIND", "0.0.0.0")
        self._token = token_digest(token)
        self._tid = None                 # transfer currently being assembled
        self.named = list(weights.named_tensors()) if hasattr(weights, "named_tensors") else list(weights)
        self.nbytes = flat_state_nbytes(self.named)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        self.view = memoryview(self.buf.numpy())
        self.num_streams, self.port = num_streams, port
        self._received = 0
        self._lock = threading.Lock()
        self._done = threading.Event()
        self._socks = []
        for i in range(num_streams):
            s = socket.socket()
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            s.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, SOCK_BUF)
            s.bind((bind_host, port + i))
            s.listen(4)
            self._socks.append(s)
            threading.Thread(target=self._serve, args=(s,), daemon=True).start()

    def _serve(self, srv: socket.socket) -> None:
        while True:
            try:
                c, _ = srv.accept()
            except OSError:
                return
            try:
                hdr = b""
                while len(hdr) < HEADER.size:
                    part = c.recv(HEADER.size - len(hdr))
                    if not part:
                        raise ConnectionError("sender closed before the header")
                    hdr += part
                tok, tid, total, off, ln = HEADER.unpack(hdr)
                if tok != self._token:
                    raise PermissionError("weight stream with a wrong token rejected")
                if total != self.nbytes or off + ln > self.nbytes:
                    raise ValueError(f"weight stream out of bounds: total={total} off={off} len={ln} expect {self.nbytes}")
                with self._lock:
                    if tid != self._tid:          # first stream of a NEW transfer: forget whatever a failed one left behind
                        self._tid, self._received = tid, 0
                        self._done.clear()
                got = 0
                while got < ln:
                    n = c.recv_into(self.view[off + got: off + ln], min(ln - got, 16 << 20))
                    if n == 0:
                        raise ConnectionError("sender closed early")
                    got += n
                c.sendall(b"OK")
                with self._lock:
                    if tid == self._tid:          # a stream of a superseded transfer does not count
                        self._received += ln
                        if self._received >= self.nbytes:
                            self._done.set()
            except Exception as e:
                log.warning("weight stream on port %s failed: %r", srv.getsockname()[1], e)
                try:
                    c.sendall(b"NO")
                except OSError:
                    pass
            finally:
                c.close()

    def wait_and_apply(self, timeout: float = 3600.0) -> bool:
        if not self._done.wait(timeout):
            with self._lock:                  # a sender that died mid-transfer must not poison the byte count of the next attempt
                self._received, self._tid = 0, None
                self._done.clear()
            raise TimeoutError("weight transfer did not complete")
        off = 0
        for _, t in self.named:
            n = t.numel() * t.element_size()
            t.view(torch.uint8).reshape(-1).copy_(self.buf[off:off + n], non_blocking=True)
            off += n
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        with self._lock:
            self._received, self._tid = 0, None
            self._done.clear()
        return True

    def close(self) -> None:
        for s in self._socks:
            s.close()
