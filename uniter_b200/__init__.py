"""uniter_b200 — B200-native (sm_100a) encoder hot path of ChenRocks/UNITER behind the
reference's own `UniterModel` contract.  See DESIGN.md / INTEGRATION.md."""
from .model import (UniterConfig, UniterModel, UniterPreTrainedModel,  # noqa: F401
                    register_lengths)
