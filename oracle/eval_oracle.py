"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement of the RetargetVid evaluator (retargetvid_eval.py:10-27,
133-194, 235-246): inclusive-pixel IoU per frame, `statistics.mean` (an exactly
rounded mean) per video and per annotator.
"""
import statistics


def bb_intersection_over_union(boxA, boxB):
	"""retargetvid_eval.py:10-27 (duplicate at smartVidCrop.py:927-944)."""
	xA = max(boxA[0], boxB[0])
	yA = max(boxA[1], boxB[1])
	xB = min(boxA[2], boxB[2])
	yB = min(boxA[3], boxB[3])
	interArea = max(0, xB - xA + 1) * max(0, yB - yA + 1)
	boxAArea = (boxA[2] - boxA[0] + 1) * (boxA[3] - boxA[1] + 1)
	boxBArea = (boxB[2] - boxB[0] + 1) * (boxB[3] - boxB[1] + 1)
	return interArea / float(boxAArea + boxBArea - interArea)


def clamp0(bb):
	"""retargetvid_eval.py:183-190"""
	return [v if v > 0 else 0 for v in bb]


def video_iou(method_bbs, annot_bbs, frame_count):
	"""retargetvid_eval.py:161-193 for one (video, annotator): frames
	0..frame_count-1, stop at the first frame either side lacks, exact mean."""
	ious = []
	for f in range(frame_count):
		if f >= len(annot_bbs) or f >= len(method_bbs):
			break
		ious.append(bb_intersection_over_union(clamp0(annot_bbs[f]), clamp0(method_bbs[f])))
	return statistics.mean(ious)


def evaluate_run(method, annots, frame_counts, ars=('1-3', '3-1')):
	"""method[ar][vid] -> list of boxes; annots[user][ar][vid] -> list of boxes.
	Returns {ar: {'per_user': [...], 'worst','best','mean' (x100)}}
	(retargetvid_eval.py:139-194, 240-246)."""
	out = {}
	for ar in ars:
		per_user_vids = [[] for _ in annots]
		for vid in sorted(method[ar].keys()):
			for u in range(len(annots)):
				per_user_vids[u].append(video_iou(method[ar][vid], annots[u][ar][vid], frame_counts[vid]))
		users = [statistics.mean(v) for v in per_user_vids]
		out[ar] = dict(per_user=users, per_user_vids=per_user_vids,
					worst=min(users) * 100, best=max(users) * 100,
					mean=statistics.mean(users) * 100)
	return out
