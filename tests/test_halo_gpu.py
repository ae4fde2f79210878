"""-m gpu: fino_halo_exchange (the row-parallel VAE's halo rows through peer mailboxes), every rank played on ONE GPU:
three "ranks" with their own frames, mailboxes and streams, several exchanges in a row (slot parity, monotonic flags)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("t,hl,w,c", [(1, 3, 6, 8), (3, 2, 10, 16), (4, 5, 40, 64)])
def test_halo_exchange_three_ranks_on_one_gpu(t, hl, w, c):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200 import ops

    world = 3
    slot = ((t * w * c * 2 + 15) // 16) * 16
    mail = [torch.zeros(ops.HALO_DATA_OFF + 4 * slot, dtype=torch.uint8, device="cuda") for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator().manual_seed(0)
    frames = [torch.zeros(t, hl + 2, w, c, dtype=torch.bfloat16, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    for seq in range(1, 6):
        bands = [torch.randn(t, hl, w, c, generator=g).bfloat16() for _ in range(world)]
        for r in range(world):
            frames[r][:, 1:1 + hl] = bands[r].cuda()
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ops.halo_exchange(frames[r], mail[r].data_ptr(), mail[r - 1].data_ptr() if r > 0 else None,
                                  mail[r + 1].data_ptr() if r < world - 1 else None, seq, slot, r)
        torch.cuda.synchronize()
        for r in range(world):
            got = frames[r].cpu()
            assert torch.equal(got[:, 1:1 + hl], bands[r])  # the band is untouched
            if r > 0:
                assert torch.equal(got[:, 0], bands[r - 1][:, hl - 1]), (seq, r, "top halo")
            else:
                assert not got[:, 0].any()  # image border: stays zero (the convolution's padding)
            if r < world - 1:
                assert torch.equal(got[:, hl + 1], bands[r + 1][:, 0]), (seq, r, "bottom halo")
            else:
                assert not got[:, hl + 1].any()
    words = [m[:32].view(torch.int32).cpu().tolist() for m in mail]
    assert words[1][0] == 5 and words[1][1] == 5 and words[0][1] == 5 and words[2][0] == 5  # flags at the last seq
    assert all(wd[2] == 0 and wd[4] == 0 and wd[5] == 0 for wd in words)  # arrival counter reset, no time-outs


def test_halo_exchange_rejects_bad_arguments():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200 import _lib, ops

    frames = torch.zeros(2, 4, 6, 8, dtype=torch.bfloat16, device="cuda")
    mail = torch.zeros(ops.HALO_DATA_OFF + 4 * 64, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.FinoError, match="mailbox slot"):
        ops.halo_exchange(frames, mail.data_ptr(), mail.data_ptr(), None, 1, 64, 0)  # 2 x 96 bytes of rows > 64
    ops.halo_exchange(frames, mail.data_ptr(), None, None, 1, 1024, 0)  # no neighbours: nothing to do
    torch.cuda.synchronize()
