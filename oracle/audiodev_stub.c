/* oracle/audiodev_stub.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for the reference's player/audiodev.c (interface at
 * player/audiodev.h:23-30), which needs ALSA headers that are not installed.
 * With -m / -o the reference CLI never opens the device (saugns.c:493-507),
 * so refusing to open is enough to link `saugns_ref`.
 */
#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>

struct SGS_AudioDev;

struct SGS_AudioDev *SGS_open_AudioDev(uint16_t channels, uint32_t *srate) {
	(void)channels; (void)srate;
	return NULL;
}
void SGS_close_AudioDev(struct SGS_AudioDev *o) { (void)o; }
uint32_t SGS_AudioDev_get_srate(const struct SGS_AudioDev *o) { (void)o; return 0; }
bool SGS_AudioDev_write(struct SGS_AudioDev *o, const int16_t *buf, uint32_t samples) {
	(void)o; (void)buf; (void)samples;
	return false;
}
