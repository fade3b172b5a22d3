"""Host-side helpers of the hot path (per-component set-up, K-sized arithmetic)."""
