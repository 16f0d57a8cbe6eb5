"""L0 foundation: ids, paths, JSON persistence, address discovery, system metrics."""
from .ids import gen_salt, hash_password, new_id, now_ms, os_name, sha256_hex  # noqa: F401
from .metrics import get_gpu_usage, get_gpu_usages, get_system_metrics, set_throughput_source  # noqa: F401
from .net import get_lan_ip, get_public_ip, is_colab, offline  # noqa: F401
from .paths import bee2bee_home, data_file, load_json, save_json  # noqa: F401
from .tracing import TRACER, Tracer  # noqa: F401
