import os, sys, torch
sys.path.insert(0, os.getcwd())
from tests import test_cnn_gpu as T
for shape in [(256, 256, 3, 1, 2, 9, 14), (512, 512, 3, 1, 1, 60, 90)]:
    cin, cout, k, stride, b, h, w = shape
    torch.manual_seed(cin + cout + k)
    conv = torch.nn.Conv2d(cin, cout, k, stride, k // 2).cuda()
    x = (torch.randn(b, cin, h, w, device='cuda') * 1.5).relu()
    with torch.no_grad():
        ref = conv(x)
    out, stats = T.run_conv_fp4(x, conv, 32)
    print(os.environ.get('CROSSLOC_B200_KW_SHARE'), shape, 'rel err %.3e' % T.rel_l2(out, ref), 'corr', float((out * ref).sum() / (ref * ref).sum()))
    # which taps are right?  one-hot filters
    for tap in range(9):
        with torch.no_grad():
            conv.weight.zero_(); conv.bias.zero_()
            conv.weight[:, :, tap // 3, tap % 3] = torch.eye(cout, cin, device='cuda')
            ref = conv(x)
        out, _ = T.run_conv_fp4(x, conv, 32)
        print('   tap', tap, 'rel err %.3e' % T.rel_l2(out, ref))
    break
