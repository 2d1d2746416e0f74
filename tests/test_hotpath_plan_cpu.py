"""Host logic of the prepared pipeline, checked without a GPU: minibatch slicing (incl. the trailing short one),
launch schedules, staging placement and the pointer arithmetic of the launch tables.  Nothing is launched -- and
trying to is an error, not a fallback."""
import pytest

from xagents_b200 import _ffi
from xagents_b200.hotpath import PPOHotPath


def addr(v):
    return v if isinstance(v, int) else (v.value or 0)


def plan(T, E, **kw):
    return PPOHotPath(T, E, (84, 84, 4), 6, device='cpu', **kw).prepare()


def test_slices_offsets_and_moment_table_cover_every_sample_once():
    hp = plan(7, 3, ppo_epochs=2, mini_batches=4)                 # N = 21, B = 5 -> 4 full + 1 short minibatch per epoch
    assert hp.N == 21 and hp.B == 5 and hp.slices == [(0, 5), (5, 10), (10, 15), (15, 20), (20, 21)]
    assert hp.n_mb == 10 and list(hp._offsets) == [0, 5, 10, 15, 20, 21, 26, 31, 36, 41, 42]
    assert hp.mb_rows == [5, 5, 5, 5, 1] * 2 and sum(hp.group_rows) == 2 * 21
    assert [a.n for a in hp._loss_args] == hp.mb_rows
    with pytest.raises(AssertionError, match='Invalid batch size to mini-batch size ratio'):
        PPOHotPath(1, 2, (4,), 2, device='cpu', mini_batches=4)


@pytest.mark.parametrize('kw,want', [({}, [8, 4, 3, 1]), ({'gather_chunk': 4}, [4, 4, 4, 4]), ({'gather_chunk': 1}, [1] * 16),
                                     ({'gather_chunk': 16}, [16]), ({'gather_chunk': [3, 1, 2, 10]}, [3, 1, 2, 10]),
                                     ({'gather_chunk': 5}, [5, 5, 5, 1])])
def test_gather_schedules(kw, want):
    hp = plan(16, 8, **kw)
    assert hp.group_sizes == want and sum(hp.group_sizes) == hp.n_mb == 16
    assert hp.kernel_launches_per_step == 2 + 16 + len(want)
    assert hp.cap == max(hp.group_rows) and hp.staging == min(2, len(want))
    with pytest.raises(AssertionError):
        plan(16, 8, gather_chunk=[8, 4])                            # does not cover the 16 minibatches


def test_launch_tables_point_into_the_right_buffers():
    T, E = 16, 8
    hp = plan(T, E, fuse_fields=False, gather_chunk=[8, 4, 3, 1])
    N, B, F = hp.N, hp.B, hp.row_bytes
    assert F == 84 * 84 * 4
    perms, obs_dst = hp.perms.data_ptr(), hp.mb_obs.data_ptr()
    for g, (fn, args) in enumerate(hp._gathers):
        first = hp.group_first[g]
        assert addr(args[0]) == hp.obs.data_ptr() and addr(args[7]) == perms + 4 * first * B       # idx of the group's first minibatch
        assert addr(args[1]) == obs_dst + (g % hp.staging) * hp.cap * F                            # staging slot
        assert args[2] == F and args[3] == N and args[6] == 4 and args[8] == hp.group_rows[g] and (args[9], args[10]) == (T, E)
    for mb, a in enumerate(hp._loss_args):
        g, slot, row0 = hp._mb_place[mb]
        assert g == max(i for i, f in enumerate(hp.group_first) if f <= mb) and row0 == (mb - hp.group_first[g]) * B
        assert a.out_scalars == hp.scalars.data_ptr() + 16 * mb
        assert a.actor_out == hp.actor_out.data_ptr() + 4 * mb * B * hp.A and a.values == hp.critic_out.data_ptr() + 4 * mb * B
        assert a.moments == hp.moments.data_ptr() + 8 * 4 * mb and a.n_moment_parts == 1
        fields = hp.mb_fields.data_ptr() + slot * 4 * 4 * hp.cap + 4 * row0                        # materialised scalar fields
        assert (a.actions, a.returns, a.old_values, a.old_log_probs) == tuple(fields + j * 4 * hp.cap for j in range(4))
        assert not a.idx
    fused = plan(T, E)                                                                             # default: read through the permutation
    for mb, a in enumerate(fused._loss_args):
        assert a.idx == fused.perms.data_ptr() + 4 * mb * B and (a.n_steps, a.n_envs) == (T, E)
        assert a.returns == fused.returns.data_ptr() and a.old_values == fused.values.data_ptr()
    assert all(args[6] == 0 for _, args in fused._gathers)                                         # no scalar fields ride along


def test_algorithmic_bytes_match_the_survey_table():
    hp = plan(128, 256)                                                                            # C3
    b = hp.algorithmic_bytes()
    N, E, K, A, F = 32768, 256, 4, 6, 28224
    assert b['gae'] == 16 * N + 4 * E and b['gather'] == K * (2 * F + 36) * N
    assert b['moments'] == K * 8 * N and b['loss'] == K * (8 * A + 24) * N
    assert b['total'] == 226272 * N + 4 * E                                                       # 226 272 B per env-step
    assert hp.h2d_bytes() == N * F + 5 * 4 * N + 4 * E + 4 * E


def test_there_is_no_cpu_path():
    hp = plan(4, 2, mini_batches=2, ppo_epochs=1)
    with pytest.raises(_ffi.XAError):
        hp.run()


def test_pipeline_binds_to_the_callers_rollout_buffers_in_place():
    """`buffers=` / `bind`: the agents hand their own `ro_*` tensors over, nothing is copied; shapes and dtypes are checked."""
    import pytest
    import torch
    T, E = 6, 4
    mine = {'obs': torch.zeros((T, E, 5), dtype=torch.float32), 'rewards': torch.zeros((T, E)), 'dones': torch.zeros((T + 1, E)),
            'actions': torch.zeros((T, E, 3)), 'returns': torch.zeros((T, E))}
    hp = PPOHotPath(T, E, (5,), 3, device='cpu', mini_batches=4, actor_kind='normal', buffers=mine)
    for name, t in mine.items():
        assert getattr(hp, name).data_ptr() == t.data_ptr()
    assert hp.values.shape == (T, E) and hp.obs.dtype == torch.float32 and hp.sync == 'event'       # no CUDA device: event sync
    other = torch.zeros((T, E))
    hp.prepare()
    hp.bind(rewards=other)
    assert hp.rewards.data_ptr() == other.data_ptr() and hp._calls is None                          # launches are re-resolved
    with pytest.raises(AssertionError, match='time-major'):
        hp.bind(dones=torch.zeros((T, E)))
    with pytest.raises(AssertionError, match='unknown rollout field'):
        hp.bind(nonsense=other)
    with pytest.raises(AssertionError, match='vector actions need fuse_fields'):
        PPOHotPath(T, E, (5,), 3, device='cpu', mini_batches=4, fuse_fields=False, buffers={'actions': mine['actions']})
    with pytest.raises(AssertionError, match='progress sync needs'):
        PPOHotPath(T, E, (5,), 3, device='cpu', mini_batches=4, sync='progress')
