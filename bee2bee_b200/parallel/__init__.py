"""NVLink mesh (peer memory, flags, topology) and multi-process launch helpers."""
