"""Host-side runtime of the partitioning classes (cache, base classes, device slabs)."""
