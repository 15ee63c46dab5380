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
