-- Optional hosted registry for a bee2bee_b200 mesh (PostgreSQL / Supabase REST).
-- `bee2bee_b200/registry.py` upserts one row per node into `active_nodes` (POST <url>/rest/v1/active_nodes with
-- `Prefer: resolution=merge-duplicates`); the reference's web gateway additionally appends token counters to
-- `messages` and reads the `system_stats` view (see SURVEY.md C19/C23).  The in-tree gateway (`gateway.py`) keeps its
-- counters in a local JSON file and does not need any of this.  Written for this repo; column set = the payload keys
-- the clients send.

create table if not exists active_nodes (
    peer_id     text primary key,
    addr        text        not null,               -- ws://host:port the node announces
    region      text        default 'Auto',
    tag         text,                               -- backend tag: hf | ollama | hf_remote | cli-<network>
    models      text[]      default '{}',
    latency_ms  double precision,
    metrics     jsonb       default '{}'::jsonb,    -- {throughput, memory_percent, gpu_percent, trust_score, ...}
    last_seen   timestamptz not null default now()
);
create index if not exists active_nodes_last_seen_idx on active_nodes (last_seen desc);

create table if not exists messages (
    id          bigserial primary key,
    node_id     text        not null,               -- peer id, or 'GLOBAL_METRICS' for mesh-wide counters
    role        text        not null default 'assistant',
    content     text,
    tokens      integer     not null default 0,
    created_at  timestamptz not null default now()
);
create index if not exists messages_node_idx on messages (node_id, created_at desc);

create table if not exists node_logs (
    id          bigserial primary key,
    peer_id     text references active_nodes (peer_id) on delete cascade,
    level       text        default 'info',
    message     text,
    created_at  timestamptz not null default now()
);

create or replace view system_stats as
select coalesce(sum(tokens), 0)                                   as total_tokens,
       count(*) filter (where content is distinct from '[Metric Log]' or tokens > 0) as total_chats,
       (select count(*) from active_nodes where last_seen > now() - interval '5 minutes') as total_users
from messages;

-- nodes that stopped syncing disappear from discovery after ten minutes
create or replace function prune_stale_nodes() returns void language sql as $$
    delete from active_nodes where last_seen < now() - interval '10 minutes';
$$;
