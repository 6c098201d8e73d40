/* log.h — level-gated stderr logging with the reference's line format "<domain>: <message>"
 * (log.h:26-43 there). */
#pragma once
#include <stdbool.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#ifndef LOG_DOMAIN
# define LOG_DOMAIN "(missing log domain, file: " __BASE_FILE__ ") "     /* the same default as the reference, log.h:3-5 */
#endif

#define LOG_EMERG 0
#define LOG_ALERT 1
#define LOG_CRIT 2
#define LOG_ERR 3
#define LOG_WARNING 4
#define LOG_NOTICE 5
#define LOG_INFO 6
#define LOG_DEBUG 7

extern int global_log_level;
void ns_host_log(int level, const char *domain, const char *fmt, ...) __attribute__((format(printf, 3, 4)));

#define log_debug(...) ns_host_log(LOG_DEBUG, LOG_DOMAIN, __VA_ARGS__)
#define log_info(...) ns_host_log(LOG_INFO, LOG_DOMAIN, __VA_ARGS__)
#define log_warning(...) ns_host_log(LOG_WARNING, LOG_DOMAIN, __VA_ARGS__)
#define log_err(...) ns_host_log(LOG_ERR, LOG_DOMAIN, __VA_ARGS__)
