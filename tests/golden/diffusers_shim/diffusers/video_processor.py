"""upstream diffusers.video_processor.VideoProcessor, the two methods the Wan FrameINO pipeline calls (recalled)."""
import torch


class VideoProcessor:
    def __init__(self, do_resize=True, vae_scale_factor=8, resample="lanczos"):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        # VaeImageProcessor.preprocess for a tensor input: [C, H, W] -> [1, C, H, W]; values in [0, 1] are normalised
        # to [-1, 1] unless the tensor already holds negative values (then upstream warns and leaves it alone)
        if image.dim() == 3:
            image = image.unsqueeze(0)
        if image.min() < 0:
            return image
        return 2.0 * image - 1.0

    def postprocess_video(self, video, output_type="np"):
        out = []
        for b in range(video.shape[0]):
            frames = video[b].permute(1, 0, 2, 3)  # [F, C, H, W]
            frames = (frames / 2 + 0.5).clamp(0, 1)  # denormalize
            out.append(frames)
        out = torch.stack(out)
        if output_type == "pt":
            return out
        if output_type == "np":
            return out.permute(0, 1, 3, 4, 2).float().cpu().numpy()
        raise ValueError(output_type)
