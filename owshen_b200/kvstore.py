"""A key-value store shaped like the reference's `trait KvStore`
(/root/reference/src/db/mod.rs:24-52: get_raw / batch_put_raw over byte blobs, `None` = delete) and its
in-memory fake `RamKvStore` (/root/reference/src/db/ram.rs:8-39), so that the GPU Merkle tree can persist its
nodes behind the same interface the node's chain state uses.  A LevelDB-backed store
(the reference's DiskKvStore, src/db/disk.rs) plugs in by implementing the same two methods."""
from typing import Dict, Iterable, Optional, Tuple


class KvStore:
    def get_raw(self, key: bytes) -> Optional[bytes]:
        raise NotImplementedError

    def batch_put_raw(self, vals: Iterable[Tuple[bytes, Optional[bytes]]]) -> None:
        raise NotImplementedError

    def put_raw(self, key: bytes, value: Optional[bytes]) -> None:      # `put` in the reference: a batch of one
        self.batch_put_raw([(key, value)])


class RamKvStore(KvStore):
    def __init__(self):
        self.db: Dict[bytes, bytes] = {}

    def get_raw(self, key: bytes) -> Optional[bytes]:
        return self.db.get(key)

    def batch_put_raw(self, vals):
        for k, v in vals:
            if v is None:
                self.db.pop(k, None)
            else:
                self.db[k] = v


class MirrorKvStore(KvStore):
    """Write overlay over a base store, the shape of the reference's MirrorKvStore
    (/root/reference/src/db/mirror.rs:7-50): reads fall through to the base, writes are buffered, `rollback()` returns
    for every overwritten key the value the BASE still holds (the undo record the chain stores per block as
    `Key::Delta`, src/blockchain/mod.rs:283-286), `buffer()` hands the pending writes over to be committed."""

    def __init__(self, base: KvStore):
        self.base = base
        self.overwrite: Dict[bytes, Optional[bytes]] = {}

    def get_raw(self, key: bytes) -> Optional[bytes]:
        if key in self.overwrite:
            return self.overwrite[key]
        return self.base.get_raw(key)

    def batch_put_raw(self, vals):
        for k, v in vals:
            self.overwrite[k] = v

    def rollback(self) -> Dict[bytes, Optional[bytes]]:
        return {k: self.base.get_raw(k) for k in self.overwrite}

    def buffer(self) -> Dict[bytes, Optional[bytes]]:
        return self.overwrite
